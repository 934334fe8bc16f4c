// vgi_svo.cu — sparse voxel octree path (placeholder until the kernels land)
#include "vgi_internal.h"
