"""Summarise an `ncu --set full` report: key raw metrics + top stalled SASS lines.
usage: python profiles/extract_ncu.py gpurun_out/prof_dw1.ncu-rep [launch index in the report] > profiles/rNN_<name>_ncu_full.txt"""
import csv, io, subprocess, sys

WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep, launch=0):
    rep = [rep, "--launch-skip", str(launch), "--launch-count", "1"]
    rows = list(csv.reader(io.StringIO(run(rep + ["--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("# %s (launch %d of the report)" % (rep[0], launch))
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print("%-68s %s %s" % (w, vals[i], units[i]))
    src = list(csv.reader(io.StringIO(run(rep[:1] + ["--page", "source", "--csv"]))))
    starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"] + [len(src)]     # one section per captured launch
    src = src[starts[launch]:starts[launch + 1]]
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in src[2:] if len(r) == len(hdr)]
    stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(float(r[ix['# Samples']] or 0) for r in data)
    agg = sorted(((sum(float(r[ix[c]] or 0) for r in data), c) for c in stall), reverse=True)
    print("\nwarp-stall samples: %d; by reason: %s" % (tot, ", ".join("%s %.0f%%" % (c[6:], 100 * v / tot) for v, c in agg[:6])))
    print("top stalled SASS instructions:")
    for r in sorted(data, key=lambda r: -float(r[ix['# Samples']] or 0))[:12]:
        rs = sorted(((float(r[ix[c]] or 0), c) for c in stall), reverse=True)[0]
        print("  %5.1f%%  %-60s %s" % (100 * float(r[ix['# Samples']] or 0) / tot, r[ix['Source']][:60], rs[1][6:]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
