"""gcc recipe for the C oracle (test infrastructure): oracle/kivi_oracle.c -> oracle/_build/libkivi_oracle.so"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "kivi_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libkivi_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force) and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=gnu11",
           "-o", SO + ".tmp", SRC, "-lm"]
    subprocess.check_call(cmd)
    os.replace(SO + ".tmp", SO)
    return SO


if __name__ == "__main__":
    print(build(force=True))
