"""Recipe for oracle/_ref/: byte-compile the UNMODIFIED reference script where it lies, outputs only into oracle/_ref/.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is mounted)

The reference's hot path is one pure-Python file, /root/reference/retrieval_data_annotation.py; "building" it is
py_compile.  The result, oracle/_ref/retrieval_data_annotation.bytecode, is a build output (git-ignored, travels to the GPU
box with the snapshot like the repo's own .so files); no reference SOURCE is copied into the repo.  bench.py's
`--impl reference` arm and `cpu_baseline` leg import `occurrence_matrix` from it unchanged (oracle/ref_loader.py) and
report kind = "reference"; tests use it to pin the restatement in oracle/jaccard_oracle.py.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under rag4dyg_b200/ may import this.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("R4D_REFERENCE_ROOT", "/root/reference")
FILES = ["retrieval_data_annotation.py", "get_train_query_time.py"]


def build(verbose=False):
    """Returns the list of compiled files (empty when the reference tree is not mounted)."""
    out_dir = os.path.join(HERE, "_ref")
    done = []
    for name in FILES:
        src = os.path.join(REF_ROOT, name)
        if not os.path.exists(src):
            continue
        os.makedirs(out_dir, exist_ok=True)
        dst = os.path.join(out_dir, name[:-3] + ".bytecode")   # not "*.pyc": snapshot tools tend to drop those
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, doraise=True)
        done.append(dst)
        if verbose:
            print("compiled", src, "->", dst)
    return done


if __name__ == "__main__":
    if not build(verbose=True):
        print(f"{REF_ROOT} not mounted: nothing to build", file=sys.stderr)
