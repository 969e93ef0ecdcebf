"""TEST INFRASTRUCTURE - builds the UNMODIFIED reference Cauchy extension (extensions/cauchy/{cauchy.cpp,cauchy_cuda.cu})
for sm_100a so tools/bench_cauchy.py can race it against libdwb on the GPU box.

    python oracle/build_ref_cauchy.py          # in the build container (needs /root/reference)

Sources are copied where they lie to the git-ignored baseline/_ref/cauchy/ (they travel to the GPU box with the
snapshot, never into history) and compiled there with torch.utils.cpp_extension into cauchy_mult.so.
Nothing in the product imports this."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/extensions/cauchy"
DST = os.path.join(ROOT, "baseline", "_ref", "cauchy")


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not mounted")
    os.makedirs(DST, exist_ok=True)
    for f in ("cauchy.cpp", "cauchy_cuda.cu", "cauchy.py", "map.h"):
        shutil.copy(os.path.join(REF, f), os.path.join(DST, f))
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils import cpp_extension
    cpp_extension.load(name="cauchy_mult", sources=[os.path.join(DST, "cauchy.cpp"), os.path.join(DST, "cauchy_cuda.cu")],
                       build_directory=DST, extra_cflags=["-O3"],
                       extra_cuda_cflags=["-O3", "-lineinfo", "--use_fast_math", "--expt-relaxed-constexpr",
                                          "-gencode", "arch=compute_100a,code=sm_100a"],
                       verbose=True)
    print("built", os.path.join(DST, "cauchy_mult.so"))


if __name__ == "__main__":
    main()
