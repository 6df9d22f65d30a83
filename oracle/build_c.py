"""Build the plain-C BoundaryMaxPooling restatement (oracle/bmp_oracle.c) into oracle/_ref/libbmp_oracle.so with gcc.
TEST INFRASTRUCTURE: loaded by tests/ only."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "libbmp_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "bmp_oracle.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run([gcc, "-O2", "-std=c99", "-shared", "-fPIC", "-o", OUT, src], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
