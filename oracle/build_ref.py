"""Build recipe for oracle/_ref -- TEST / BASELINE INFRASTRUCTURE.

The reference (pliang279/factorized) is pure Python: "compiling" its model file means byte-compiling it.  This script
byte-compiles the UNMODIFIED /root/reference/mfm_model.py, where it lies, into oracle/_ref/mfm_model.pyc.bin (the .bin suffix keeps
snapshot tools that drop *.pyc from dropping it; git-ignored,
NOT gpurun-ignored: it travels to the GPU box like the built .so).  No reference source enters the repository.
`bench.py --impl reference` then times the reference's own MFM class (kind "reference"); without the artefact it falls
back to the oracle port (kind "port").  __graft_entry__.build() calls this when /root/reference is present.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/mfm_model.py"
OUT = os.path.join(HERE, "_ref", "mfm_model.pyc.bin")


def build(verbose=True):
    if not os.path.exists(REF_SRC):
        if verbose:
            print("oracle/build_ref: %s not present (GPU box / no reference tree): nothing to do" % REF_SRC)
        return None
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    py_compile.compile(REF_SRC, cfile=OUT, doraise=True)
    if verbose:
        print("oracle/build_ref: %s -> %s" % (REF_SRC, OUT))
    return OUT


if __name__ == "__main__":
    sys.exit(0 if build() or not os.path.exists(REF_SRC) else 1)
