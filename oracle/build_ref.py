"""TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Recipe that stages the UNMODIFIED reference for the CPU arm of bench.py.

    python oracle/build_ref.py            # /root/reference/{rgcn,src,hyperbolic_src}/*.py -> oracle/_ref/ (byte copies)

oracle/_ref/ is git-ignored (reference sources never enter the history) but travels to the GPU box with the snapshot,
exactly like the built libregcn_b200.so: `bench.py --impl reference` and the `cpu_baseline` leg import the reference's
own src/main.py test() loop from there under the DGL stand-in of oracle/fake_dgl.py.  Nothing else may import it.
A manifest with the sha256 of every staged file is written next to the copies (oracle/_ref/MANIFEST.json).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REGCN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("rgcn", "src", "hyperbolic_src")


def build(quiet=False):
    if not os.path.isdir(REF):
        if not quiet:
            print(f"{REF} not present: keeping whatever oracle/_ref/ already holds")
        return os.path.isdir(OUT)
    manifest = {}
    for pkg in PACKAGES:
        src_dir, dst_dir = os.path.join(REF, pkg), os.path.join(OUT, pkg)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith(".py"):
                continue
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            with open(os.path.join(dst_dir, name), "rb") as f:
                manifest[f"{pkg}/{name}"] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(manifest)} reference files under {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
