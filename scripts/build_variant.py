"""Builds a variant of libmamcts_b200.so with extra -D flags for one source (A/B experiments; search.cu unless
VARIANT_SRC names another):
    python scripts/build_variant.py NAME -DMACRO=VALUE ...   ->  scripts/micro/libmamcts_NAME.so
Run it with MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_NAME.so."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
name, flags = sys.argv[1], sys.argv[2:]
SRC = os.environ.get("VARIANT_SRC", "search.cu")
inc = ["-I", os.path.join(ROOT, "include"), "-I", g.CSRC] + g._nccl_include()
obj = f"/tmp/search_{name}.o"
subprocess.check_call([g._nvcc()] + g.NVCC_FLAGS + flags + inc + ["-c", os.path.join(g.CSRC, SRC), "-o", obj])
objs = [os.path.join(g.CSRC, s.replace(".cu", ".o")) for s in g.CUDA_SOURCES if s != SRC] + [obj]
out = os.path.join(ROOT, "scripts", "micro", f"libmamcts_{name}.so")
subprocess.check_call([g._nvcc(), "-Wno-deprecated-gpu-targets", "-shared", "-o", out] + objs + ["-lcudart", "-ldl", "-lpthread"])
print(out)
