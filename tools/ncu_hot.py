"""Developer aid: summarise an .ncu-rep (one kernel): duration, DRAM bytes, issue-active, and the SASS lines with the
most stall samples together with their dominant stall reason.   python tools/ncu_hot.py rep.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
            "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:70s} {v} {u}")
    src = page(rep, "source")
    h = src[1]
    data = src[2:]
    ia, isamp = h.index("Instructions Executed"), h.index("# Samples")
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = collections.Counter()
    for r in data:
        for c in cols:
            tot[h[c]] += int(r[c] or 0)
    n = sum(int(r[isamp]) for r in data)
    print("samples", n, "warp instructions", sum(int(r[ia]) for r in data))
    print("stalls:", ", ".join(f"{k[6:]} {v} ({100 * v / max(n, 1):.0f}%)" for k, v in tot.most_common(8)))
    top = sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:top_n]
    for i, r in sorted(top):
        why = max(cols, key=lambda c: int(r[c] or 0))
        print(f"{i:5d} {r[1].strip()[:64]:64s} exec {r[ia]:>9s} samples {r[isamp]:>5s} ({100 * int(r[isamp]) / max(n, 1):4.1f}%) {h[why][6:]}")

main()
