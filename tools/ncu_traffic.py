"""From an `ncu --page raw --csv` export of the C5 operators: per-kernel duration, DRAM bytes and pipe
utilisation -> profiles/ncu_traffic_r01.json (bench.py reads `dram_bytes` for roofline.traffic) and a
text summary.  usage: python tools/ncu_traffic.py <raw.csv> <out.json> <out.txt> <tag>"""
import csv, json, re, sys

LABELS = [(r"rowfft_kernel", "edfdv.row"), (r"rowfft2_kernel", "edfdv.row2"), (r"pass13_kernel<\d+, 0, 0", "vdfdx.pass1"), (r"pass2_kernel<\d+, 0", "vdfdx.pass2"),
          (r"pass13_kernel<\d+, 0, 1", "vdfdx.pass3"), (r"fp_reg_kernel", "fp_step"), (r"Xmodes2Prog|XmodesProg", "xmodes"),
          (r"fp_kernel", "fp_step(shared-memory kernel)"),
          (r"midfft_kernel<.*, 1, \d+>", "edfdv.mid"), (r"midfft_kernel<.*, 0, \d+>", "vdfdx.mid"), (r"moments_warp_kernel", "moments")]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    raw, out_json, out_txt, tag = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res, lines = {}, ["ncu --set full --clock-control none, 16384x16384 fp64, %s" % tag, ""]
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        label = next((lab for pat, lab in LABELS if re.search(pat, name)), None)
        if label is None or label in res:
            continue
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
        res[label] = {"kernel": name, "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "ncu_ms": dur_ms}
        fk = "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"
        if fk in ix and r[ix[fk]]:
            res[label]["fp64_pipe_pct"] = float(r[ix[fk]].replace(",", ""))
        lines.append("== %s   (%s)" % (name, label))
        for w in WANT:
            if w in ix:
                lines.append("%-86s %s %s" % (w, r[ix[w]], units[ix[w]]))
        for h in hdr:
            if "stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                v = float(r[ix[h]] or 0)
                if v >= 0.3:
                    lines.append("%-86s %.3f" % (h, v))
        lines.append("")
    json.dump({"source": tag, "kernels": res}, open(out_json, "w"), indent=1)
    open(out_txt, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
