"""Build the CUDA library in-tree (sm_100a only):  python -m gpuintegration_b200.build [--oracle]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build_library(jobs=8):
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}"])
    return os.path.join(HERE, "libpagani_b200.so")


def build_oracle(jobs=8):
    """Test infrastructure: CPU restatement + (when /root/reference exists) the reference harness."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), f"-j{jobs}", "all"])


if __name__ == "__main__":
    print(build_library())
    if "--oracle" in sys.argv:
        build_oracle()
