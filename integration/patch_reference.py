#!/usr/bin/env python
"""Builds the PATCHED include overlay a Neon maintainer's change amounts to, without copying reference
sources into this repository: the files below are read from the reference tree, edited in memory and written to an
overlay directory that is put in FRONT of the reference's include paths (oracle/Makefile.ref, target ref_b200).

  benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h
      the body of LbmContainers::iteration (:283-325) is replaced by integration/lbm_iteration_b200.inc — a device-managed
      container that calls nlbm_d3q19_*_dense_step — and integration/lbm_shim_prelude.inc is added after the includes.
  benchmarks/lbm-lid-driven-cavity-flow/src/LbmIteration.h
      copied verbatim next to it (it includes "LbmTools.h" by a quoted name, which is looked up next to the including file).
No library header is touched (the reference's Container::factoryDeviceManaged does not compile when instantiated; the shim
brings its own managed container, integration/B200ManagedContainer.h).

    python integration/patch_reference.py /root/reference oracle/_ref/patched
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main(ref: str, out: str) -> None:
    # ---- LbmTools.h
    src = open(os.path.join(ref, "benchmarks/lbm-lid-driven-cavity-flow/src/LbmTools.h")).read()
    begin = src.index("    static auto\n    iteration(Neon::set::StencilSemantic stencilSemantic,")
    end = src.index("#define COMPUTE_MASK_WALL")
    body = open(os.path.join(HERE, "lbm_iteration_b200.inc")).read()
    prelude = open(os.path.join(HERE, "lbm_shim_prelude.inc")).read()
    patched = src[:begin] + body + "\n" + src[end:]
    last_inc = [m for m in re.finditer(r'^#include .*$', patched, re.M)][-1]
    patched = patched[:last_inc.end()] + "\n\n" + prelude + patched[last_inc.end():]
    os.makedirs(os.path.join(out, "bench"), exist_ok=True)
    open(os.path.join(out, "bench", "LbmTools.h"), "w").write(patched)
    # LbmIteration.h includes "LbmTools.h" by a quoted name, which is looked up next to the including file first: an
    # unmodified copy of it in the overlay makes that lookup find the patched LbmTools.h
    li = open(os.path.join(ref, "benchmarks/lbm-lid-driven-cavity-flow/src/LbmIteration.h")).read()
    open(os.path.join(out, "bench", "LbmIteration.h"), "w").write(li)
    print(f"patched overlay written to {out}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
