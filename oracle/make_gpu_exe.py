#!/usr/bin/env python
"""make_gpu_exe.py -- TEST INFRASTRUCTURE: build the reference's own program with the B200 matcher plugged in.

Copies /root/reference/Core/src to a scratch directory (never into this repository), applies the edits
INTEGRATION.md describes, drops in patternmatching_b200/ref_glue/mpgpu.{c,h}, and links the result against
patternmatching_b200/libpm_b200.so.  Outputs (git-ignored, shipped to the GPU box):

  oracle/_ref/exe_gpu       every reference algorithm (AC, LMAC, MPBG) + the two GPU rows, driver unchanged
                            except for the batched call at measure.c:292-294; STREAM_BUFFER_SIZE as shipped (100 KiB)
  oracle/_ref/exe_gpu_big   AC + the two GPU rows only (LMAC 0.5 MB/s and MPBG 265 B/s cannot read a large stream),
                            STREAM_BUFFER_SIZE = 16 MiB with the three chunk arrays made static (they are stack
                            arrays in the reference, measure.c:243-245) and page-locked once by the glue

Edits are made by anchor replacement; every anchor must occur exactly once, so a reference that has changed makes
this script fail instead of mis-patching.
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def edit(path, anchor, repl):
    with open(path) as f:
        text = f.read()
    if text.count(anchor) != 1:
        sys.exit(f"make_gpu_exe: anchor {anchor!r} occurs {text.count(anchor)} times in {path}")
    with open(path, "w") as f:
        f.write(text.replace(anchor, repl))


def build(ref, out, big):
    src = os.path.join(ref, "Core", "src")
    tmp = tempfile.mkdtemp(prefix="pm_gpu_exe_")
    try:
        for f in glob.glob(os.path.join(src, "*.[ch]")):
            shutil.copy(f, tmp)
        for f in ("mpgpu.c", "mpgpu.h"):
            shutil.copy(os.path.join(ROOT, "patternmatching_b200", "ref_glue", f), tmp)
        p = lambda name: os.path.join(tmp, name)
        # 1. Core/src/mps.h:20-25 -- the enum
        edit(p("mps.h"), "\tMPS_SIZE\n", "\tMPS_GPU,      // B200 exact dictionary scan\n\tMPS_GPU_KR,   // B200 Karp-Rabin stages\n\tMPS_SIZE\n")
        # 2. Core/src/mps.c:17-20, 120-124 -- include + registration
        edit(p("mps.c"), '#include "mplmac.h"\n', '#include "mplmac.h"\n#include "mpgpu.h"\n')
        edit(p("mps.c"), "\tmps_lmac_register();\n", "\tmps_lmac_register();\n\tmps_gpu_register();\n\tmps_gpu_kr_register();\n")
        # 3. Core/src/measure.c:292-294 -- the hot loop
        edit(p("measure.c"), '#include "conf.h"\n', '#include "conf.h"\n#include "mpgpu.h"\n')
        edit(p("measure.c"),
             "\t\t\tfor (j = 0; j < len_read; ++j) {\n\t\t\t\talgo_results[j] = read_char_func(obj, stream_buffer[j]);\n\t\t\t}\n",
             "\t\t\tif (mps_read_block_of(inst->algo)) {\n"
             "\t\t\t\tmps_read_block_of(inst->algo)(obj, stream_buffer, len_read, algo_results);\n"
             "\t\t\t} else for (j = 0; j < len_read; ++j) {\n\t\t\t\talgo_results[j] = read_char_func(obj, stream_buffer[j]);\n\t\t\t}\n")
        units = sorted(glob.glob(os.path.join(tmp, "*.c")))
        if big:
            edit(p("measure.c"), "#define STREAM_BUFFER_SIZE (100 * 1024)", "#define STREAM_BUFFER_SIZE (16 * 1024 * 1024)")
            edit(p("measure.c"), "\tchar stream_buffer[STREAM_BUFFER_SIZE];", "\tstatic char stream_buffer[STREAM_BUFFER_SIZE];")
            edit(p("measure.c"), "\tpattern_id_t algo_results[STREAM_BUFFER_SIZE];", "\tstatic pattern_id_t algo_results[STREAM_BUFFER_SIZE];")
            edit(p("measure.c"), "\tpattern_id_t real_results[STREAM_BUFFER_SIZE];", "\tstatic pattern_id_t real_results[STREAM_BUFFER_SIZE];")
            edit(p("mps.h"), "\tMPS_LMAC,     // Multi-Pattern Low-Memory Aho-Corasick\n", "")
            edit(p("mps.h"), "\tMPS_BG,       // Multi-Pattern Brausler-Galil\n", "")
            edit(p("mps.c"), "\tmps_bg_register();\n", "")
            edit(p("mps.c"), "\tmps_lmac_register();\n", "")
            units = [u for u in units if os.path.basename(u) not in ("mpbg.c", "mplmac.c")]
        os.makedirs(out, exist_ok=True)
        exe = os.path.join(out, "exe_gpu_big" if big else "exe_gpu")
        cmd = ["gcc", "-O2", "-w", "-U_FORTIFY_SOURCE", "-I" + tmp, "-I" + os.path.join(ROOT, "include")] + \
              (["-DMPGPU_REGISTER_BUFFERS"] if big else []) + units + \
              ["-L" + os.path.join(ROOT, "patternmatching_b200"), "-lpm_b200",
               "-Wl,-rpath,$ORIGIN/../../patternmatching_b200", "-o", exe]
        subprocess.check_call(cmd)
        return exe
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref"))
    a = ap.parse_args()
    if not os.path.exists(os.path.join(a.ref, "Core", "src", "mps.c")):
        print(f"make_gpu_exe: {a.ref} not present -- using prebuilt oracle/_ref/exe_gpu* (if any)")
        return 0
    for big in (False, True):
        print("built", build(a.ref, a.out, big))
    return 0


if __name__ == "__main__":
    sys.exit(main())
