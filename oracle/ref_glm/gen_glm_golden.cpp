// gen_glm_golden.cpp — TEST INFRASTRUCTURE. Generates tests/golden/glm_golden.json from the
// REFERENCE'S OWN vendored glm (compiled in place from /root/reference/Dependencies, nothing is
// copied into this repo), with the reference's build flags (ref: VFS/pch.h:33-35). It replays the
// reference host's glm call sequences that feed the hot path's UBOs:
//   Camera::updateCamera            ref: VFS/Camera.cpp:110-121 (defaults Camera.h:48-56)
//   DirectionalLight::setTransform  ref: VFS/DirectionalLight.cpp:21-47
//   Voxelizer::setViewProjection    ref: VFS/RenderPass/Clipmap/Voxelizer.cpp:298-327
// Build + run: oracle/ref_glm/gen_golden.sh (needs /root/reference; not run on the GPU box).
#define GLM_FORCE_SIZE_T_LENGTH
#define GLM_FORCE_DEPTH_ZERO_TO_ONE
#define GLM_FORCE_LEFT_HANDED
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>

static void put_mat(FILE* f, const char* key, const glm::mat4& m, bool last = false)
{
    // bit patterns (u32) so the comparison is exact
    const float* p = &m[0][0];
    fprintf(f, "    \"%s\": [", key);
    for (int i = 0; i < 16; ++i) {
        uint32_t u;
        memcpy(&u, p + i, 4);
        fprintf(f, "%u%s", u, i < 15 ? ", " : "");
    }
    fprintf(f, "]%s\n", last ? "" : ",");
}

static void camera_case(FILE* f, const char* name, glm::vec3 pos, glm::vec3 dir, float aspect, bool last)
{
    const glm::vec3 up(0.0f, -1.0f, 0.0f);
    const float fovy = 60.0f;
    glm::mat4 view = glm::lookAt(pos, pos + dir, up);
    glm::mat4 proj = glm::perspective(glm::radians(fovy), aspect, 0.01f, 5000.0f);
    fprintf(f, "  \"%s\": {\n", name);
    fprintf(f, "    \"position\": [%.9g, %.9g, %.9g], \"direction\": [%.9g, %.9g, %.9g], \"aspect\": %.9g,\n",
            pos.x, pos.y, pos.z, dir.x, dir.y, dir.z, aspect);
    put_mat(f, "view_proj", proj * view);
    put_mat(f, "view_proj_inv", glm::inverse(view) * glm::inverse(proj), true);
    fprintf(f, "  }%s\n", last ? "" : ",");
}

static void light_case(FILE* f, const char* name, glm::vec3 origin, glm::vec3 direction, bool last)
{
    const float zNear = 0.1f, zFar = 30.0f; // DirectionalLight.h:50-51
    glm::vec3 d = glm::normalize(direction);
    glm::mat4 view = glm::lookAt(origin, origin + d, glm::vec3(0.0f, 1.0f, 0.0f));
    glm::mat4 proj = glm::ortho(-16.0f, 16.0f, -16.0f, 16.0f, zNear, zFar);
    fprintf(f, "  \"%s\": {\n", name);
    fprintf(f, "    \"origin\": [%.9g, %.9g, %.9g], \"direction\": [%.9g, %.9g, %.9g],\n", origin.x, origin.y,
            origin.z, direction.x, direction.y, direction.z);
    uint32_t u[3];
    memcpy(u, &d[0], 12);
    fprintf(f, "    \"direction_normalized\": [%u, %u, %u],\n", u[0], u[1], u[2]);
    put_mat(f, "view", view);
    put_mat(f, "proj", proj, true);
    fprintf(f, "  }%s\n", last ? "" : ",");
}

static void voxelizer_case(FILE* f, const char* name, glm::ivec3 minCorner, uint32_t R, float voxelSize, bool last)
{
    const glm::vec3 regionGlobal = glm::vec3(glm::uvec3(R)) * voxelSize;
    const glm::vec3 minCornerGlobal = glm::vec3(minCorner) * voxelSize;
    const glm::vec3 eye = minCornerGlobal + glm::vec3(0.0f, 0.0f, regionGlobal.z);
    glm::mat4 vp[3];
    vp[0] = glm::ortho(-regionGlobal.z, regionGlobal.z, -regionGlobal.y, regionGlobal.y, 0.1f, regionGlobal.x) *
            glm::lookAt(eye, eye + glm::vec3(1.0f, 0.0f, 0.0f), glm::vec3(0.0f, 1.0f, 0.0f));
    vp[1] = glm::ortho(-regionGlobal.x, regionGlobal.x, -regionGlobal.z, regionGlobal.z, 0.1f, regionGlobal.y) *
            glm::lookAt(eye, eye + glm::vec3(0.0f, 1.0f, 0.0f), glm::vec3(0.0f, 0.0f, -1.0f));
    vp[2] = glm::ortho(-regionGlobal.x, regionGlobal.x, -regionGlobal.y, regionGlobal.y, 0.1f, regionGlobal.z) *
            glm::lookAt(minCornerGlobal, minCornerGlobal + glm::vec3(0.0f, 0.0f, 1.0f), glm::vec3(0.0f, 1.0f, 0.0f));
    fprintf(f, "  \"%s\": {\n", name);
    fprintf(f, "    \"min_corner\": [%d, %d, %d], \"resolution\": %u, \"voxel_size\": %.9g,\n", minCorner.x, minCorner.y,
            minCorner.z, R, voxelSize);
    put_mat(f, "view_proj_x", vp[0]);
    put_mat(f, "view_proj_y", vp[1]);
    put_mat(f, "view_proj_z", vp[2], true);
    fprintf(f, "  }%s\n", last ? "" : ",");
}

int main(int argc, char** argv)
{
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    if (!f) return 1;
    fprintf(f, "{\n");
    fprintf(f, "  \"_generator\": \"oracle/ref_glm/gen_glm_golden.cpp against the reference's vendored glm (GLM_VERSION %d), LH + ZO; matrices are column-major float32 bit patterns\",\n", GLM_VERSION);
    camera_case(f, "camera_cfg1_cornell", glm::vec3(0.0f), glm::vec3(0.0f, 0.0f, -1.0f), 1.0f, false);
    camera_case(f, "camera_cfg2_atrium", glm::vec3(-8.0f, 3.0f, 0.0f), glm::vec3(1.0f, 0.0f, 0.0f), 1920.0f / 1080.0f, false);
    camera_case(f, "camera_oblique", glm::vec3(2.5f, 4.0f, -7.25f), glm::normalize(glm::vec3(-0.3f, -0.2f, 0.9f)), 1.5f, false);
    light_case(f, "light_reference_default", glm::vec3(0.0f, 30.0f, -5.3f), glm::vec3(0.0f, -1.0f, 0.2f), false);
    light_case(f, "light_cfg1_cornell", glm::vec3(0.0f, 20.0f, -3.5f), glm::vec3(0.0f, -1.0f, 0.2f), false);
    voxelizer_case(f, "voxelizer_level0_r128", glm::ivec3(-64), 128, 0.125f, false);
    voxelizer_case(f, "voxelizer_level3_r256_moved", glm::ivec3(-144, -122, -128), 256, 0.5f, true);
    fprintf(f, "}\n");
    if (f != stdout) fclose(f);
    return 0;
}
