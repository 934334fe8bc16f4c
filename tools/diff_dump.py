"""Compare two vgi dumps (vk_voxel_cone_tracing_b200/dump.py): python tools/diff_dump.py a.npz b.npz — exit status 1 when they differ."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vk_voxel_cone_tracing_b200 import dump  # noqa: E402

if __name__ == "__main__":
    lines = dump.diff(dump.load(sys.argv[1]), dump.load(sys.argv[2]))
    print("\n".join(lines) if lines else "identical")
    sys.exit(1 if lines else 0)
