"""profiles/ncu_kernel_table.json from `ncu --page raw --csv` dumps: per workload and pass label the DRAM bytes and warp instructions of ONE launch of its
kernel (bench.py reads `roofline.traffic` / `issue_slots` from it and names the csv in `traffic_source`).

    python tools/ncu_table.py <workload> profiles/<raw.csv> [<workload> <raw.csv> ...]

When a kernel appears several times in a dump the launch with the longest duration is taken (validation frames do more work)."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_TO_PASS = {
    "k_rtdgi_reproject": "rtdgi reproject", "k_extract_half_fused": "extract half-res inputs", "k_rtdgi_validate": "rtdgi validate", "k_rtdgi_trace": "rtdgi trace",
    "k_rtdgi_validity_integrate": "validity integrate", "k_rtdgi_restir_temporal": "restir temporal", "k_rtdgi_restir_spatial": "restir spatial", "k_rtdgi_restir_resolve": "restir resolve",
    "k_rtdgi_temporal": "rtdgi temporal", "k_rtdgi_spatial": "rtdgi spatial", "k_reprojection_map": "reprojection map",
    "k_rtr_trace": "reflection trace", "k_rtr_validate": "reflection validate", "k_rtr_restir_temporal": "rtr restir temporal", "k_rtr_resolve": "reflection resolve",
    "k_rtr_temporal": "reflection temporal", "k_rtr_cleanup": "reflection cleanup",
    "k_taa_reproject": "reproject taa", "k_taa_filter_input_tiled": "taa filter input", "k_taa_filter_history_tiled": "taa filter history", "k_taa_filter_history": "taa filter history",
    "k_taa_input_prob": "taa input prob", "k_taa_prob_filter": "taa prob filter", "k_taa_prob_filter2": "taa prob filter2", "k_taa_tiled": "taa", "k_taa_tiled_upsampling": "taa",
    "k_ircache_validate": "ircache validate", "k_ircache_trace": "ircache trace", "k_ircache_trace_access": "ircache trace access", "k_ircache_sum": "ircache sum",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def table_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        label = KERNEL_TO_PASS.get(name)
        if not label:
            continue
        dur = num(r[idx["gpu__time_duration.sum"]]) or 0.0
        rd, wr = num(r[idx["dram__bytes_read.sum"]]) or 0.0, num(r[idx["dram__bytes_write.sum"]]) or 0.0
        rd *= scale.get(units[idx["dram__bytes_read.sum"]], 1.0); wr *= scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        e = {"dram_bytes": rd + wr, "warp_inst": num(r[idx["smsp__inst_executed.sum"]]), "ncu_duration_ms": dur * (1e-3 if units[idx["gpu__time_duration.sum"]] == "us" else 1.0),
             "registers": num(r[idx["launch__registers_per_thread"]]), "occupancy_pct": num(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
             "issue_active_pct": num(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]) if "smsp__issue_active.avg.pct_of_peak_sustained_active" in idx else None,
             "threads_per_inst": num(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]), "source": os.path.relpath(path, ROOT)}
        if label not in out or e["ncu_duration_ms"] > out[label]["ncu_duration_ms"]:
            out[label] = e
    return out


if __name__ == "__main__":
    dst = os.path.join(ROOT, "profiles", "ncu_kernel_table.json")
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    a = sys.argv[1:]
    for wl, path in zip(a[0::2], a[1::2]):
        table.setdefault(wl, {}).update(table_of(path))
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
    for wl in table:
        print(wl, {k: (round(v["dram_bytes"] / 1e6, 1), v["warp_inst"]) for k, v in table[wl].items()})
