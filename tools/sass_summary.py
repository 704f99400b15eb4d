"""Static SASS instruction counts of the hot kernels in the built library (cuobjdump -sass), as a markdown table.

Usage: python tools/sass_summary.py [round-tag] > profiles/<round>_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pypolymlp_b200", "lib", "libpolymlp_b200.so")
OPS = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "REDG", "ATOMG", "ATOMS", "LDS", "LDG", "STG", "DFMA", "DMUL", "DADD", "BAR"]
HOT = ["k_syrk_sk2", "k_syrk_sk(", "k_lrows_v4n<3, 8>", "k_lrows_v4<3, 8>", "k_lrows_v4a<3, 8>", "k_xrows_v6", "k_pair_anlm<4, true>",
       "k_features_v3<4, 3, 256>", "k_lrows_big<12, 512>", "k_lrows_big<8, 256>", "k_xrows_v5", "k_lrows_v3<3, 8, false>",
       "k_features_v4r<1, 512, 5>", "k_anlm_eval<4, true>", "k_eval_features_lb<5>", "k_eval_features_la<3>", "k_eval_pairs_rc<4, 12, false>",
       "k_eval_features<4, 3>", "k_eval_pairs_v2", "k_neighbor_cl_count", "k_neighbor_cl_fill", "k_neighbor_mask<false>",
       "k_pack_upper", "k_xe_reduce"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, fp_atom, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            counts[cur][m.group(1)] += 1
            if m.group(1) in ("REDG", "ATOMG", "RED", "ATOM") and ".F64" in m.group(2):
                fp_atom[cur] += 1
    names = demangle(list(counts))
    short = {k: re.sub(r"\(.*", "(", v.replace("(anonymous namespace)::", "")).replace("void ", "") for k, v in names.items()}
    print(f"# Round {tag[1:].lstrip('0')} — SASS summary of `pypolymlp_b200/lib/libpolymlp_b200.so` (`cuobjdump -sass`, static instruction "
          "counts; `python tools/sass_summary.py`)\n")
    print("tcgen05 / UMMA has no f64 kind: the fp64 tensor path on sm_100a is `mma.sync.m8n8k4.f64` = `DMMA.8x8x4`.  `UBLKCP` = 1-D TMA "
          "bulk copy (`cp.async.bulk`), `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`, `REDG` / `ATOMG` = global atomics "
          "(`f64 atomics` = those of them that add doubles).\n")
    print("| kernel | " + " | ".join(OPS) + " | f64 atomics |")
    print("|---|" + "---|" * (len(OPS) + 1))
    for h in HOT:
        for k, s in short.items():
            if h in s:
                label = s.rstrip("(")
                print(f"| `{label}` | " + " | ".join(str(counts[k][o]) for o in OPS) + f" | {fp_atom[k]} |")
                break
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"| **whole library** ({len(counts)} kernels) | " + " | ".join(str(tot[o]) for o in OPS) + f" | {sum(fp_atom.values())} |")


if __name__ == "__main__":
    main()
