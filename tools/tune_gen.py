#!/usr/bin/env python
"""Plan tuner: generates, builds (here) and runs (on the GPU box) many compile-time plan variants of one length.

    python tools/tune_gen.py build <spec.json>      # -> tools/_build/tune_<name> (nvcc, sm_100a), no GPU needed
    tools/_build/tune_<name> [batch] [iters]        # on the GPU box; prints one JSON line per variant

spec.json: {"name": "f32_1024", "type": "float", "mode": "C2C", "nfft": 1024, "radices": [4,4,4,4,4],
            "variants": [{"groups": [2,2,1], "team": 64, "tpc": 4, "logpad": 4, "minblocks": 3}, ...]}
Development aid only; the chosen plans are written into kissfft_b200/csrc/kf_plan_list.h by hand.
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tools", "_build")
TYPEFLAGS = {"float": "-Dkiss_fft_scalar=float", "double": "-Dkiss_fft_scalar=double -DKF_IS_DOUBLE", "int16_t": "-DFIXED_POINT=16",
             "int32_t": "-DFIXED_POINT=32"}
NVCC = ["nvcc", "-std=c++20", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
        "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]

HEAD = r'''
#include "%(root)s/include/kiss_fft.h"
#define KF_SCALAR_BYTES ((int)sizeof(kiss_fft_scalar))
#include "%(root)s/kissfft_b200/csrc/kf_kernels.cuh"
#include "%(root)s/kissfft_b200/csrc/kf_plan_list.h"
#include "%(root)s/tools/tune_common.h"
using namespace kf;
'''


def lst(x):
    return "{" + ", ".join(str(int(v)) for v in x) + "}"


def gen(spec):
    os.makedirs(BUILD, exist_ok=True)
    name = spec["name"]
    variants = spec["variants"]
    nper = 6
    files = []
    for ci in range(0, len(variants), nper):
        chunk = variants[ci:ci + nper]
        src = HEAD % {"root": ROOT}
        for j, v in enumerate(chunk):
            idx = ci + j
            src += "struct V%d { static constexpr PlanDesc D = make_plan(%d, %s, %s, %d, %d, %d, %d, %d, %d, %d, %d, %d, %d); };\n" % (
                idx, spec["nfft"], lst(v.get("radices", spec["radices"])), lst(v["groups"]), v["team"], v["tpc"], v["logpad"], v["minblocks"], v.get("nstage", 0), v.get("nbuf", 2), v.get("twmode", 0), v.get("shfl", 0), v.get("paired", 0), v.get("hoist", 0))
        src += "void register_chunk_%d(std::vector<TuneEntry>& out) {\n" % (ci // nper)
        for j, v in enumerate(chunk):
            idx = ci + j
            label = "r%s_g%s_t%d_c%d_p%d_b%d_s%d_n%d_w%d_x%d_q%d%s" % ("".join(map(str, v.get("radices", spec["radices"]))), "".join(map(str, v["groups"])), v["team"], v["tpc"], v["logpad"], v["minblocks"], v.get("nstage", 0), v.get("nbuf", 2), v.get("twmode", 0), v.get("shfl", 0), v.get("paired", 0), "_h%d" % v["hoist"] if v.get("hoist", 0) else "")
            src += '    out.push_back(make_entry<V%d, k%s>("%s"));\n' % (idx, spec["mode"], label)
        src += "}\n"
        path = os.path.join(BUILD, "tune_%s_%d.cu" % (name, ci // nper))
        open(path, "w").write(src)
        files.append(path)
    nchunks = len(files)
    main = HEAD % {"root": ROOT}
    for c in range(nchunks):
        main += "void register_chunk_%d(std::vector<TuneEntry>& out);\n" % c
    main += "int main(int argc, char** argv) {\n    std::vector<TuneEntry> v;\n"
    for c in range(nchunks):
        main += "    register_chunk_%d(v);\n" % c
    main += "    return tune_main(argc, argv, v, %d, k%s);\n}\n" % (spec["nfft"], spec["mode"])
    mpath = os.path.join(BUILD, "tune_%s_main.cu" % name)
    open(mpath, "w").write(main)
    files.append(mpath)
    return files


def build(spec):
    files = gen(spec)
    tf = TYPEFLAGS[spec["type"]]
    objs = [f[:-3] + ".o" for f in files]

    def cc(fo):
        f, o = fo
        r = subprocess.run(NVCC + tf.split() + ["-c", f, "-o", o], capture_output=True, text=True)
        if r.returncode:
            print(r.stdout[-3000:], r.stderr[-3000:])
            raise SystemExit("compile failed: " + f)
        return r.stderr

    with ThreadPoolExecutor(8) as ex:
        logs = list(ex.map(cc, zip(files, objs)))
    open(os.path.join(BUILD, "tune_%s.ptxas.log" % spec["name"]), "w").write("\n".join(logs))
    exe = os.path.join(BUILD, "tune_" + spec["name"])
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", *objs, "-o", exe], check=True)
    for f in files + objs:
        os.unlink(f)
    print(exe, len(spec["variants"]), "variants")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        for p in sys.argv[2:]:
            build(json.load(open(p)))
