"""Build recipe for the checker libraries.  TEST INFRASTRUCTURE ONLY.

  oracle/liboracle.so           CPU restatement (plenvdb_oracle.cpp), self-contained, builds anywhere.
  oracle/_ref/libref_host.so    harness over the reference's NanoVDB headers, host only      } only when
  oracle/_ref/libref_gpu.so     + the reference's unmodified .cu kernels for sm_100a          } /root/reference
  oracle/_ref/render_utils_ref.so  the reference's torch extension `render_utils_cuda`        } is present
  oracle/_ref/adam_upd_ref.so      and `adam_upd_cuda`, compiled for sm_100a                  }

The reference sources are compiled where they lie (include / source paths into /root/reference); nothing
is copied into the repository.  The torch extensions do not compile against torch 2.11 as shipped
(`AT_DISPATCH_FLOATING_TYPES(x.type(), ...)`): the recipe pipes them through a one-token
`sed 's/.type()/.scalar_type()/'` into a temporary file under the git-ignored oracle/_ref/ and deletes it
after compiling.  oracle/_ref/ is git-ignored but not gpurun-ignored, so the built libraries travel to
the GPU box.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLENVDB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd, **kw):
    res = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if res.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout[-4000:], res.stderr[-4000:]))
    return res


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_oracle(force=False):
    src = os.path.join(HERE, "plenvdb_oracle.cpp")
    out = os.path.join(HERE, "liboracle.so")
    if force or _stale(out, [src, os.path.join(HERE, "plenvdb_oracle.h")]):
        _run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


def have_reference():
    return os.path.isdir(os.path.join(REF, "plenvdb", "lib", "vdb"))


def build_ref_host(force=False):
    src = os.path.join(HERE, "ref_harness.cu")
    out = os.path.join(OUT, "libref_host.so")
    if force or _stale(out, [src]):
        os.makedirs(OUT, exist_ok=True)
        _run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-x", "c++",
              "-I", os.path.join(REF, "openvdb", "nanovdb"), "-o", out, src])
    return out


def build_ref_gpu(force=False):
    src = os.path.join(HERE, "ref_harness.cu")
    out = os.path.join(OUT, "libref_gpu.so")
    vdb = os.path.join(REF, "plenvdb", "lib", "vdb")
    if force or _stale(out, [src]):
        os.makedirs(OUT, exist_ok=True)
        inc = ["-I", os.path.join(REF, "openvdb", "nanovdb"), "-I", vdb]
        common = [NVCC] + ARCH + ["-O3", "-std=c++17", "--extended-lambda", "-Xcompiler", "-fPIC", "-w"] + inc
        objs = []
        for name in ["plenvdb", "densityvdb", "colorvdb", "renderer"]:
            obj = os.path.join(OUT, "ref_%s.o" % name)
            extra = ["-include", "thrust/execution_policy.h"] if name == "renderer" else []
            _run(common + extra + ["-c", os.path.join(vdb, name + ".cu"), "-o", obj])
            objs.append(obj)
        hobj = os.path.join(OUT, "ref_harness.o")
        _run(common + ["-DREF_WITH_CUDA", "-c", src, "-o", hobj])
        _run([NVCC] + ARCH + ["-shared", "-o", out] + objs + [hobj, "-lcublas", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
        for o in objs + [hobj]:
            os.remove(o)
    return out


def _build_torch_ext(name, cpp, cu, force=False):
    """Compile one of the reference's torch extensions into oracle/_ref/<name>.so (a CPython module)."""
    out = os.path.join(OUT, name + ".so")
    cuda_dir = os.path.join(REF, "plenvdb", "lib", "cuda")
    if not (force or _stale(out, [os.path.join(cuda_dir, cpp), os.path.join(cuda_dir, cu)])):
        return out
    os.makedirs(OUT, exist_ok=True)
    import torch
    from torch.utils import cpp_extension as ce
    incs = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"]]:
        incs += ["-I", p]
    defs = ["-DTORCH_EXTENSION_NAME=" + name, "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    tmp_cu = os.path.join(OUT, "_tmp_%s.cu" % name)
    tmp_cpp = os.path.join(OUT, "_tmp_%s.cpp" % name)
    try:
        with open(tmp_cu, "w") as f:
            subprocess.run(["sed", "s/\\.type()/.scalar_type()/g", os.path.join(cuda_dir, cu)], stdout=f, check=True)
        with open(tmp_cpp, "w") as f:
            subprocess.run(["sed", "s/\\.type()/.scalar_type()/g", os.path.join(cuda_dir, cpp)], stdout=f, check=True)
        o_cu, o_cpp = tmp_cu + ".o", tmp_cpp + ".o"
        _run([NVCC] + ARCH + ["-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-w"] + defs + incs + ["-c", tmp_cu, "-o", o_cu])
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-w"] + defs + incs + ["-I", "/usr/local/cuda/include", "-c", tmp_cpp, "-o", o_cpp])
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        _run(["g++", "-shared", "-o", out, o_cu, o_cpp, "-L", tlib, "-Wl,-rpath," + tlib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda",
              "-ltorch_python", "-lc10", "-lc10_cuda", "-L", "/usr/local/cuda/lib64", "-lcudart"])
    finally:
        for p in (tmp_cu, tmp_cpp, tmp_cu + ".o", tmp_cpp + ".o"):
            if os.path.exists(p):
                os.remove(p)
    return out


def build_ref_torch_exts(force=False):
    return [_build_torch_ext("render_utils_ref", "render_utils.cpp", "render_utils_kernel.cu", force),
            _build_torch_ext("adam_upd_ref", "adam_upd.cpp", "adam_upd_kernel.cu", force),
            _build_torch_ext("total_variation_ref", "total_variation.cpp", "total_variation_kernel.cu", force)]


def build_ref_callers(force=False):
    """The reference's own Python CALLERS of the boundary — plenvdb/lib/grid.py (QueryVerticalInVDB, VDBGrid) and
    plenvdb/lib/masked_adam.py (VDBAdam) — byte-compiled from where they lie into oracle/_ref/*.pycode (CPython's .pyc format), like the C++ / CUDA
    reference is compiled into oracle/_ref/*.so: no source enters the repo, and the compiled modules travel to the GPU box,
    where tests/test_reference_callers_gpu.py runs them on top of plenvdb_b200 (the drop-in exercised by the reference's code)."""
    import py_compile
    outs = []
    for name in ("grid", "masked_adam"):
        src = os.path.join(REF, "plenvdb", "lib", name + ".py")
        out = os.path.join(OUT, "ref_caller_%s.pycode" % name)      # a .pyc under another suffix: snapshots of the tree drop *.pyc
        if force or _stale(out, [src]):
            os.makedirs(OUT, exist_ok=True)
            py_compile.compile(src, cfile=out, dfile="reference:plenvdb/lib/%s.py" % name, doraise=True)
        outs.append(out)
    return outs


def build_all(force=False, with_torch_exts=True):
    built = [build_oracle(force)]
    if have_reference():
        built.append(build_ref_host(force))
        built.append(build_ref_gpu(force))
        if with_torch_exts:
            built += build_ref_torch_exts(force)
        built += build_ref_callers(force)
    return built


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, with_torch_exts="--no-torch-exts" not in sys.argv):
        print(p)
