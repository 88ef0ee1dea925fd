"""Markdown for DESIGN.md §6 from the bench lines kept under profiles/ (so that the document quotes files, not memory).
   python tools/design_results.py <n1 bench json> [<reference json>] [<nK bench json> ...]"""
import json, sys


def load(p):
    return json.loads(open(p).read().strip().splitlines()[-1])


def main(paths):
    b = load(paths[0])
    print(f"Headline (`{paths[0]}`, driver command `python bench.py --gpus 1 --steps {b['steps']} --warmup {b['warmup']}`, clocks {b['clocks']['sm_mhz']} MHz, reasons {b['clocks']['reasons']}):\n")
    print("| configuration (BASELINE `configs[i]`) | ms/frame (inputs in HBM) | ms/frame e2e (host in/out) | GI rays/s | CPU oracle, all host threads (ms/frame) |")
    print("|---|---|---|---|---|")
    for e in b["configs"]:
        c = e["config"]; cb = e.get("cpu_baseline", {})
        cpu = f"{cb['ms_per_step']:.0f}" + (" (¼ width and height)" if "1/4" in cb.get("sample", "") else "") if cb else "—"
        print(f"| [{c['baseline_config']}] `{c['workload']}` {c['resolution'][0]}×{c['resolution'][1]}" + (f" → {c['output_resolution'][0]}×{c['output_resolution'][1]}" if c['output_resolution'] != c['resolution'] else "") +
              f" | **{e['ms_per_step']:.3f}** | {e['e2e']['ms_per_step']:.3f} | {e['value'] / 1e6:.0f} M | {cpu} |")
    r = b["roofline"]
    print(f"\nDominant kernel of the headline: `{r['kernel']}` {r['kernel_ms']:.3f} ms = {100 * r['kernel_share_of_step']:.0f} % of the frame; algorithmic {r['algorithmic_bytes_per_launch'] / 1e6:.1f} MB per launch ⇒ "
          f"**{r['achieved']:.0f} GB/s = {100 * r['frac']:.1f} % of the measured HBM peak** ({r['peak']:.0f} GB/s); ncu DRAM traffic {r['traffic'] / 1e6:.1f} MB per launch (`{r['traffic_source']}`) — no wasted re-reads. "
          f"Whole frame: {r['frame']['algorithmic_bytes'] / 1e6:.0f} MB algorithmic ⇒ {r['frame']['achieved_gbs']:.0f} GB/s = {100 * r['frame']['frac']:.1f} %. "
          + (f"Issue slots of the dominant kernel: {r['issue_slots']['warp_inst_per_launch'] / 1e6:.0f} M warp instructions per launch ⇒ {r['issue_slots']['achieved_ginst_s']:.0f} G/s = "
             f"**{100 * r['issue_slots']['frac']:.0f} % of the issue peak** ({r['issue_slots']['peak_ginst_s']:.0f} G warp-inst/s)." if r.get("issue_slots") else ""))
    pp = r["per_pass_ms"]
    print("\nPer pass (CUDA events around every pass, one queue, direct launches; ms): " + ", ".join(f"{k} {v:.3f}" for k, v in sorted(pp.items(), key=lambda kv: -kv[1])[:16]) + f"; sum of all {sum(pp.values()):.2f}.")
    if "fast_math" in b and "ms_per_step" in b["fast_math"]:
        f = b["fast_math"]
        print(f"\nPrice of the numeric contract: `libkjb_fast.so` renders the same frames in {f['ms_per_step']:.2f} ms ⇒ exact / fast = **{f['exact_over_fast']:.2f}×**.")
    k = b.get("ray_kinds") or b["configs"][0].get("ray_kinds")
    if k:
        print(f"Rays per headline frame: {k['closest_hit_per_frame'] / 1e6:.2f} M closest-hit (with shading) + {k['any_hit_per_frame'] / 1e6:.2f} M any-hit (visibility).")
    rest = paths[1:]
    for p in rest:
        d = load(p)
        if d.get("impl") == "reference":
            cb = d["cpu_baseline"]
            print(f"\nReference arm (`{p}`): the oracle on {cb['cores']} host threads, {d['steps']} timed frames of the headline workload: {d['ms_per_step']:.0f} ms/frame "
                  f"(min / median / max {cb['spread_ms']['min']:.0f} / {cb['spread_ms']['median']:.0f} / {cb['spread_ms']['max']:.0f}) = {d['value'] / 1e6:.2f} M rays/s ⇒ e2e ratio ≈ {b['e2e']['value'] / d['value']:.0f}× (context, not a result: the CPU arm moves several-fold between boxes).")
        else:
            n = d["n_gpus"]
            print(f"\nN = {n} (`{p}`): headline {d['ms_per_step']:.3f} ms/frame (speed-up {b['ms_per_step'] / d['ms_per_step']:.2f}×, efficiency {b['ms_per_step'] / d['ms_per_step'] / n:.2f}), e2e {d['e2e']['ms_per_step']:.3f} ms; parity `{d['parity']['mode'].split(' (')[0]}` ok = {d['parity']['ok']}; "
                  + "; ".join(f"`{e['config']['workload']}` {e['ms_per_step']:.3f} ms, parity {'bit-identical' if e['parity']['bands_bit_identical'] else 'statistical'} ok = {e['parity']['ok']}" for e in d["configs"][1:]) + ".")


if __name__ == "__main__":
    main(sys.argv[1:])
