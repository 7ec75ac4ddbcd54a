"""Load/store widths, shared/local memory use and register counts of the hot kernels,
from the built objects (cuobjdump -sass / -res-usage).  Written to profiles/ as the
evidence the round-1 verdict asked for (load widths, no spills).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "openpnm_b200", "build")
WANT = {
    "solver.o": ["k_dia_spmv_dotILi3ELb1E", "k_update_xrpILb0E", "k_update_xrpILb1E",
                 "k_sell_spmv_dot_wILb1ELi256ELi8ELi128E", "k_csr_stream_spmv_dotILb1E",
                 "k_dia_smoothILi3E", "k_spmv_csr_vecILi8E"],
    "amg.o": ["k_amg_spmv_f32", "k_amg_row_build", "k_amg_propose", "k_amg_restrict",
              "k_amg_prolong"],
    "p2p.o": ["k_dia_edges_p2pILi3E", "k_update_xrp_p2pILb0E"],
    "assembly.o": ["k_fill_conns", "k_emit_laplacian", "k_bc_emit", "k_conduit_conductance"],
    "topology.o": ["k_uf_union"],
}


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


def main():
    for obj, names in WANT.items():
        path = os.path.join(BUILD, obj)
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                cur = m.group(1)
            elif cur and "REG:" in line:
                usage[cur] = line.strip()
                cur = None
        blocks = re.split(r"\n\s*Function : ", sass)
        for b in blocks[1:]:
            fn = b.split("\n", 1)[0].strip()
            if not any(w in fn for w in names):
                continue
            ops = collections.Counter(re.findall(r"\b(LDG\.[A-Z0-9.]+|STG\.[A-Z0-9.]+|LDS[.A-Z0-9]*|STS[.A-Z0-9]*|"
                                                 r"LDL[.A-Z0-9]*|STL[.A-Z0-9]*|ATOMG?[.A-Z0-9]*|RED[.A-Z0-9]*|"
                                                 r"DFMA|DADD|DMUL|BAR\.SYNC[.A-Z0-9]*|SHFL[.A-Z0-9]*|NANOSLEEP)\b", b))
            print(f"== {obj}: {demangle(fn)[:150]}")
            print(f"   {usage.get(fn, '(no res-usage line)')}")
            print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())))
            spills = sum(v for k, v in ops.items() if k.startswith(("LDL", "STL")))
            print(f"   local-memory (spill) instructions: {spills}")


if __name__ == "__main__":
    main()
