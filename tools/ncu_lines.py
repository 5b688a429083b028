"""Per-source-line view of an ncu capture made with --import-source on (run here, on the CPU box).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep kernels_tc 'tc_detect_kernel<(int)4>' [top]

Joins `ncu --page source --csv` (SASS rows in address order, with stall samples and executed-instruction counts) with the
line table of the matching cubin (`nvdisasm -g`), then prints, per source line: stall samples, warp instructions executed,
and the dominant stall reasons.  The cubin comes from the in-tree libsyldet_cuda.so, so build the same sources first.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "syllable-detector-swift_b200", "libsyldet_cuda.so")


def sass_lines(cubin_stem, kernel_substr):
    """[(mnemonic, file, line)] for the kernel's instructions, in address order."""
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_stem) and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur, active = [], ("?", 0), False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            active = kernel_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((m.group(2).strip(), cur[0], cur[1]))
    return out


def main():
    rep, stem, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    mangled = sys.argv[5] if len(sys.argv) > 5 else re.sub(r"[^A-Za-z0-9_]", "", kern.split("<")[0])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern.split("<")[0]], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lines = sass_lines(stem, mangled)
    if len(lines) != len(body):
        print("warning: %d SASS rows in the report vs %d in the cubin (rebuild the same sources?)" % (len(body), len(lines)))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    total = 0
    for r, (mn, f, l) in zip(body, lines):
        s = int(r[col["# Samples"]] or 0)
        key = (f, l)
        agg[key]["samples"] += s
        agg[key]["inst"] += int(r[col["Instructions Executed"]] or 0)
        agg[key]["thread_inst"] += int(r[col["Thread Instructions Executed"]] or 0)
        for h in stall_cols:
            agg[key][h] += int(r[col[h]] or 0)
        total += s
    src_cache = {}

    def src(f, l):
        if f not in src_cache:
            p = os.path.join(ROOT, "syllable-detector-swift_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        t = src_cache[f]
        return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""

    tot_inst = sum(v["inst"] for v in agg.values())
    print("total samples %d, warp instructions %d" % (total, tot_inst))
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        reasons = sorted(((v[h], h[6:]) for h in stall_cols if v[h]), reverse=True)[:3]
        print("%5.1f%% %8d inst  %-22s %-32s %s" % (100.0 * v["samples"] / max(total, 1), v["inst"], "%s:%d" % (f, l),
                                                   " ".join("%s=%d" % (n, c) for c, n in reasons), src(f, l)))


if __name__ == "__main__":
    main()
