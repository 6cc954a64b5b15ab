"""Regenerate profiles/<tag>_sass_excerpts.txt from the built library (no GPU needed).

    python tools/make_sass_excerpts.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "griduniverse_b200", "lib", "libgu_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+", "   ", line))
names = demangle(list(funcs))


def opcode(line):
    m = re.match(r"\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    return m.group(2) if m else "?"


def find(pattern):
    for k, v in names.items():
        if re.search(pattern, v):
            return k, v
    raise KeyError(pattern)


def histogram(lines, top=28):
    c = collections.Counter(opcode(l) for l in lines)
    return " ".join("%s:%d" % kv for kv in c.most_common(top))


def section(f, title, pattern, want, n_lines=12, window=None):
    k, full = find(pattern)
    lines = funcs[k]
    f.write("\n## %s\n# %s\n# %d SASS instructions; opcode histogram:\n# %s\n" % (title, full[:150], len(lines), histogram(lines)))
    if window:                       # a contiguous excerpt around the first match of `window`
        idx = next(i for i, l in enumerate(lines) if re.search(window, l))
        f.write("# %d instructions from the first %s on (one unrolled step / cell group):\n" % (n_lines, window))
        for l in lines[idx:idx + n_lines]:
            f.write(l.rstrip() + "\n")
    else:
        f.write("# first lines with %s:\n" % " / ".join(want))
        hits = [l for l in lines if any(w in l for w in want)][:n_lines]
        for l in hits:
            f.write(l.rstrip() + "\n")


with open(os.path.join(ROOT, "profiles", tag + "_sass_excerpts.txt"), "w") as f:
    f.write("# cuobjdump -sass griduniverse_b200/lib/libgu_b200.so (sm_100a), built from the committed sources\n"
            "# (tools/make_sass_excerpts.py).  UTMALDG = cp.async.bulk.tensor (2-D TMA tiles), SYNCS = mbarrier ops,\n"
            "# FMUL2 / FADD2 / FFMA2 = packed f32x2 arithmetic (two cells per instruction).\n")
    f.write("\n## kernels containing UTMALDG (instantiations grouped by kernel and by UTMALDG count)\n")
    groups = collections.OrderedDict()
    for k, lines in funcs.items():
        n = sum("UTMALDG" in l for l in lines)
        if n:
            base = re.sub(r"<.*", "", re.sub(r"^void gu::", "", names[k]))
            groups.setdefault((base, n), []).append(names[k])
    for (base, n), members in groups.items():
        f.write("  %-28s %3d instantiations x %2d UTMALDG   e.g. %s\n" % (base, len(members), n,
                                                                         re.search(r"<.*?>\(", members[0]).group(0)[:-1][:90]))
    section(f, "rollout, cfg 4 (two envs per lane, 2-stage ring): one unrolled action row",
            r"rollout_info8_tma_kernel<2, false, true, gu::RingStd, false, false>", [], 30, window=r"LDS\.64")
    section(f, "rollout, cfg 3 (one env per lane, 4-stage ring): the step chain",
            r"rollout_info8_tma_kernel<1, false, true, gu::RingMid, false, false>", [], 28, window=r"LDS\.64")
    section(f, "rollout: TMA issue and mbarrier waits", r"rollout_info8_tma_kernel<2, false, true, gu::RingStd, false, false>",
            ["UTMALDG", "SYNCS"], 10)
    section(f, "fused-greedy fp32 sweep (LDG window, packed f32x2 arithmetic)",
            r"sweep_tiled_kernel<float, 3, false, 2, false, false>", ["FMUL2", "FADD2", "FFMA2"], 12)
    section(f, "fused-greedy fp64 sweep (TMA-staged window)",
            r"sweep_tiled_kernel<double, 3, false, 2, false, true>", ["UTMALDG", "SYNCS"], 10)
    section(f, "fused-greedy fp32 sweep, peer-memory variant",
            r"sweep_tiled_kernel<float, 3, false, 2, true, false>", ["FFMA2", "MEMBAR", "ST.E"], 10)
    section(f, "greedy extraction fp32: tie masks assembled on the FMA pipe (FSET -> FFMA into a 2^23 accumulator, PRMT)",
            r"sweep_tiled_kernel<float, 3, true, 2, false, false>", [], 24, window=r"FSET\.BF\.EQ")
    section(f, "resident look_step_ahead service: system-scope polling loop", r"look_server_kernel", ["LD.E", "ST.E", "CS2R", "MEMBAR"], 10)
print("wrote profiles/%s_sass_excerpts.txt" % tag)
