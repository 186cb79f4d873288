"""Build the TEST-ONLY host compile of csrc/k2_core.h (see k2_host.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "k2_host.cpp")
CORE = os.path.join(HERE, "..", "..", "svd_quantization_task_merging_b200", "csrc", "k2_core.h")
OUT = os.path.join(HERE, "libk2host.so")


def build() -> str:
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE))
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                    "-o", OUT, SRC, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
