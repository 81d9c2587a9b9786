#!/usr/bin/env python
"""Compiles the reference's OWN controller code against the drop-in classes: src/simulation.{h,cpp} with the three
edits of INTEGRATION.md section 3 applied to a scratch copy under /tmp (nothing of the reference enters this
repository), `-DEQ_B200_IN_EQ_TREE` so that gpuHSL / gpuFD build on the reference's real src/eQ.h, and the interface
shims of oracle/ standing in for the headers this image lacks (dolfin.h, mpi.h, petsc*.h, chipmunk.h).  Syntax and
type check only (`g++ -fsyntax-only`): every member simulation.cpp reaches into on the solver -- shell->mesh->
num_vertices(), shell->mesh_coords, shell->dof_from_vertex, D11/D22/D12 through eQ::mpi's operator>>, solution_vector
through operator<< / >>, totalBoundaryFlux, setBoundaryValues, getBoundaryFlux, writeDiffusionFiles, writeDataFiles,
topChannelData/bottomChannelData, finalize -- must exist on gpuHSL with a compatible type.

    python scripts/check_dropin_compiles.py [/root/reference]        exit 0 = compiles
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FACTORY = '''        gpuHSL::config cfg;
        cfg.boundaryType  = eQ::data::parameters["boundaryType"];
        cfg.trapType      = eQ::data::parameters["trapType"];
        cfg.lengthScaling = eQ::data::parameters["lengthScaling"];
        cfg.simulationFlowRate            = eQ::data::parameters["simulationFlowRate"];
        cfg.simulationChannelLengthLeft   = eQ::data::parameters["simulationChannelLengthLeft"];
        cfg.simulationChannelLengthRight  = eQ::data::parameters["simulationChannelLengthRight"];
        cfg.channelSolverNumberIterations = eQ::data::parameters["channelSolverNumberIterations"];
        if (cfg.boundaryType == "MICROFLUIDIC_TRAP") {
            const char *walls[4] = {"left", "right", "top", "bottom"};
            for (int w = 0; w < 4; ++w) {
                auto v = eQ::data::parameters["boundaries"][walls[w]][1].get<std::vector<double>>();
                for (int k = 0; k < 3; ++k) cfg.boundaries[w][k] = v[k];
            }
        }
        diffusionSolver = std::make_shared<gpuHSL>(cfg);
'''
# (file, old, new): the edits INTEGRATION.md section 3 lists
EDITS = [
    ("simulation.h", "std::shared_ptr<fenicsInterface>    diffusionSolver;", "std::shared_ptr<gpuHSL>             diffusionSolver;"),
    ("simulation.h", '#include "fHSL.h"', '#include "fHSL.h"\n#include "gpuHSL.h"'),
    ("simulation.cpp", "        diffusionSolver = std::make_shared<fenicsInterface>();//create instance\n", FACTORY),
    ("simulation.cpp", "diffusionSolver = std::make_shared<fenicsInterface>(fenicsParams);",
     "diffusionSolver = std::make_shared<gpuHSL>(fenicsParams);"),
]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src")):
        print(f"{ref}/src absent: nothing to check")
        return 2
    work = "/tmp/eq_b200_dropin/src"
    shutil.rmtree(os.path.dirname(work), ignore_errors=True)
    os.makedirs(work)
    for name in ("simulation.h", "simulation.cpp"):
        shutil.copy(os.path.join(ref, "src", name), work)
    for name, old, new in EDITS:
        path = os.path.join(work, name)
        s = open(path, encoding="utf-8-sig").read()
        if old not in s:
            print(f"edit does not apply to {name}: {old!r}")
            return 1
        open(path, "w").write(s.replace(old, new, 1))
    o = os.path.join(ROOT, "oracle")
    inc = ["-DEQ_B200_IN_EQ_TREE", f"-I{ROOT}/eq_b200/host", f"-I{o}/shim_dolfin", f"-I{o}/shim", f"-I{o}/shim_cpm/a/b",
           f"-I{o}/shim_petsc", f"-I{ref}/src", f"-I{ref}/src/abm"]
    rc = 0
    for src in (os.path.join(work, "simulation.cpp"), f"{ROOT}/eq_b200/host/gpuHSL.cpp", f"{ROOT}/eq_b200/host/gpuFD.cpp"):
        r = subprocess.run(["/usr/bin/g++", "-std=c++14", "-w", "-fsyntax-only"] + inc + [src], capture_output=True, text=True)
        errs = [l for l in r.stderr.splitlines() if "error" in l]
        print(("ok      " if r.returncode == 0 else "FAILED  ") + src)
        for l in errs[:20]:
            print("   ", l)
        rc |= r.returncode
    return 1 if rc else 0


if __name__ == "__main__":
    sys.exit(main())
