"""Summaries of ncu outputs: `launches <csv>` aggregates a gpu__time_duration launch list by kernel;
`stalls <ncu-rep>` prints the top stalled SASS lines; `metrics <ncu-rep>` prints the headline counters."""
import collections, csv, re, subprocess, sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.1f}  {k}")


def stalls(rep, top=25):
    top = int(top)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
    sc = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Address":
            break
        if r[i_s].isdigit():
            data.append((int(r[i_s]), r))
    tot = sum(n for n, _ in data)
    print("samples", tot)
    for n, r in sorted(data, key=lambda x: -x[0])[:top]:
        st = sorted([(int(r[i]) if r[i].isdigit() else 0, h) for i, h in sc], reverse=True)[:2]
        print(f"{n:6d} {100 * n / tot:5.1f}%  {r[i_src].strip()[:64]:64s} {st}")


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct", "launch__registers_per_thread", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__cycles_elapsed.max", "lts__t_bytes.sum ", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
            "smsp__inst_executed.sum ", "lts__t_sectors_srcunit_tex_op_read.sum ", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
    for row in r[2:]:
        print("kernel:", row[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else "")
        for i, h in enumerate(hdr):
            if any(h.startswith(k.strip()) for k in keys) and "per_second" not in h and ".pct_of_peak_sustained_elapsed" not in h[len(h) - 40:] or h in ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"):
                print(f"  {h} [{units[i]}] = {row[i]}")


if __name__ == "__main__":
    {"launches": launches, "stalls": stalls, "metrics": metrics}[sys.argv[1]](*sys.argv[2:])
