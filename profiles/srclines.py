"""Per-source-line instruction counts of the kernels in an ncu report (needs --import-source on / -lineinfo).

    python profiles/srclines.py gpurun_out/x.ncu-rep [kernel-substring] [top]
"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if want:
        cmd += ["-k", "regex:" + want]       # without a filter ncu prints only the first kernels of a big report
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    state = dict(kernel=None, data=[], ie=None, si=None)

    def flush():
        k, data = state["kernel"], state["data"]
        if k is None or not data or want not in k:
            return
        tot = sum(d[0] for d in data)
        print(f"## {k}   warp instructions executed: {tot}")
        for n, ln, src, smp in sorted(data, reverse=True)[:top]:
            print(f"{100 * n / tot:5.1f}%  smp={smp:>6}  {ln:>5}  {src.strip()[:150]}")

    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            flush()
            state.update(kernel=r[1], data=[])
        elif r[0] == "Line No":
            state["ie"], state["si"] = r.index("Instructions Executed"), r.index("# Samples")
        elif state["ie"] is not None and len(r) > state["ie"] and r[0].isdigit():
            try:
                state["data"].append((int(r[state["ie"]]), r[0], r[1], r[state["si"]]))
            except ValueError:
                pass
    flush()


if __name__ == "__main__":
    main()
