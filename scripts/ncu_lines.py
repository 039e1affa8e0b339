"""Per-source-line shares of warp-state samples and executed instructions of every kernel in an .ncu-rep that was
captured with `--set full --import-source on` (needs the sources at the paths -lineinfo recorded):

    python scripts/ncu_lines.py gpurun_out/prof_reference.ncu-rep profiles/r02_final_C5_source_lines.csv [min_pct]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    res = collections.OrderedDict()
    fn = fp = None
    ie = ns = None
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            fp = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1].split("(")[0].split("::")[-1]
        elif r[0] == "Line No":
            ie, ns = r.index("Instructions Executed"), r.index("# Samples")
        elif r[0] != "" and ie is not None:
            try:
                key = (fn, fp, int(r[0]))
            except ValueError:
                continue
            def num(v):
                try:
                    return int(v)
                except ValueError:
                    return 0
            s, i = res[key][1:] if key in res else (0, 0)
            res[key] = (r[1].strip()[:100], s + num(r[ns]), i + num(r[ie]))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "file", "line", "samples_pct", "instructions_pct", "source"])
        for kern in sorted({k[0] for k in res}):
            items = [(k, v) for k, v in res.items() if k[0] == kern]
            ts, ti = sum(v[1] for _, v in items) or 1, sum(v[2] for _, v in items) or 1
            for k, v in sorted(items, key=lambda kv: (kv[0][1], kv[0][2])):
                if 100 * v[1] / ts >= min_pct or 100 * v[2] / ti >= min_pct:
                    w.writerow([kern, k[1], k[2], f"{100 * v[1] / ts:.2f}", f"{100 * v[2] / ti:.2f}", v[0]])
    print(out)


if __name__ == "__main__":
    main()
