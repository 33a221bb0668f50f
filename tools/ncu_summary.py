"""Summarise an ncu report (.ncu-rep, captured with `--set full --import-source on`) as a small markdown file for
profiles/: duration, DRAM / L2 / shared traffic, pipe utilisation, occupancy, registers, stall reasons per issue, and
the source lines with the most stall samples.  Runs where ncu is installed (no GPU needed).

  python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" [algorithmic_bytes] [kernel-name regex] > profiles/r02_ncu_x.md
(with a kernel-name regex, the first matching launch of a multi-kernel report is summarised)
"""
import csv
import io
import subprocess
import sys


KFILTER = []


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"] + KFILTER, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    alg = float(sys.argv[3]) if len(sys.argv) > 3 and float(sys.argv[3]) > 0 else None
    if len(sys.argv) > 4:
        KFILTER.extend(["-k", "regex:" + sys.argv[4], "-c", "1"])
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

    def g(k, d=""):
        return m.get(k, (d, ""))[0]

    def f(k):
        try:
            return float(g(k, "nan").replace(",", ""))
        except ValueError:
            return float("nan")

    def scale(k):  # bytes with ncu's unit prefixes
        v, u = m.get(k, ("nan", "byte"))
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
        try:
            return float(v.replace(",", "")) * mult
        except ValueError:
            return float("nan")

    dur_v, dur_u = m.get("gpu__time_duration.sum", ("nan", "us"))
    dur_s = float(dur_v.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(dur_u, 1e-6)
    rd, wr = scale("dram__bytes_read.sum"), scale("dram__bytes_write.sum")
    print("# %s\n" % title)
    print("Kernel `%s`, grid %s x block %s, %s registers/thread, %s dynamic shared memory per block; one launch under"
          " `ncu --set full --clock-control none` (duration under the profiler, cold caches — the bench numbers are event-timed outside ncu).\n" % (
              g("Kernel Name", g("launch__function_name", "?")), g("launch__grid_size"), g("launch__block_size"), g("launch__registers_per_thread"),
              " ".join(m.get("launch__shared_mem_per_block_dynamic", ("?", "")))))
    print("| quantity | value |\n|---|---|")
    print("| duration | %.3f ms |" % (dur_s * 1e3))
    print("| DRAM read / written | %.3f GB / %.3f GB |" % (rd / 1e9, wr / 1e9))
    print("| DRAM throughput (read+write over the duration) | %.0f GB/s |" % ((rd + wr) / dur_s / 1e9))
    if alg:
        print("| algorithmic bytes of this launch (SURVEY 8d formula) | %.3f GB -> traffic / algorithmic = %.2f |" % (alg / 1e9, (rd + wr) / alg))
    for k, lab in (("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of ncu peak"),
                   ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput, % of peak"),
                   ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
                   ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts, % of peak"),
                   ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active, %"),
                   ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
                   ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy, % of max warps"),
                   ("launch__waves_per_multiprocessor", "waves per SM")):
        if k in m:
            print("| %s | %s |" % (lab, g(k)))
    print("| warp instructions executed | %s |" % g("smsp__inst_executed.sum"))
    print("\nStall reasons (warps stalled per issued instruction, > 0.2):\n")
    st = []
    for h in hdr:
        if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            v = f(h)
            if v == v and v > 0.2:
                st.append((v, h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    for v, n in sorted(st, reverse=True):
        print("* %s %.2f" % (n, v))
    # source page (SASS view): instructions with the most stall samples, and the share per opcode
    try:
        src = page(rep, "source")
        hi = next(i for i, r in enumerate(src) if "Source" in r and "# Samples" in r)
        sh = src[hi]
        si, ki = sh.index("Source"), sh.index("# Samples")
        lines, by_op = [], {}
        for r in src[hi + 1:]:
            try:
                v = float(r[ki].replace(",", ""))
            except (ValueError, IndexError):
                continue
            text = " ".join(r[si].split())
            lines.append((v, text))
            op = text.split()[1] if text.startswith("@") and len(text.split()) > 1 else text.split()[0]
            by_op[op] = by_op.get(op, 0.0) + v
        tot = sum(v for v, _ in lines) or 1.0
        print("\nStall samples by SASS opcode (share of all samples, > 2 %):\n")
        for op, v in sorted(by_op.items(), key=lambda kv: -kv[1]):
            if v / tot > 0.02:
                print("* %.1f %% `%s`" % (100.0 * v / tot, op))
        print("\nSASS instructions with the most stall samples:\n")
        for v, t in sorted(lines, reverse=True)[:10]:
            print("* %.1f %% `%s`" % (100.0 * v / tot, t[:110]))
    except Exception as e:  # the raw page is the essential part
        print("\n(source page not summarised: %r)" % (e,))


if __name__ == "__main__":
    main()
