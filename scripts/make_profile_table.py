"""profiles/*_per_kernel_table.md from a bench.py JSON line (its roofline_layers rows)."""
import json
import sys


def main(src, dst, tag):
    d = json.load(open(src))
    rows = d["roofline_layers"]
    tot = sum(r["ms"] for r in rows)
    out = []
    out.append(f"# {tag} per-kernel-class device times inside one bench step (N={d['n_gpus']}, {d['dtype']} conv operands)\n")
    out.append(f"Source: `bench.py` roofline pass (CUDA events around every C-ABI call of one extra in-line step, batch "
               f"{d['config']['frames_per_gpu']} nuScenes-shaped frames, {d['config']['voxels_per_frame']} voxels/frame).  "
               f"Timed (pipelined) step = {d['ms_per_step']:.2f} ms ({d['value']:.1f} frames/s); e2e = {d['e2e']['value']:.1f} frames/s; "
               f"sum of the classes below = {tot:.2f} ms (the geometry classes run on the input pipeline's side stream in the timed steps).")
    out.append("Algorithmic flops/bytes: SURVEY.md section 8(d) formulas with the measured N, V and pair counts. Peaks: "
               f"{d['roofline']['peak_source']}.\n")
    out.append("| kernel class | calls | ms | share | TFLOP/s | % bf16 peak | GB/s | % HBM peak |")
    out.append("|---|---|---|---|---|---|---|---|")
    for r in rows:
        tf = r.get("tflops")
        out.append("| {} | {} | {:.3f} | {:.1f}% | {} | {} | {} | {} |".format(
            r["kernel"], r["calls"], r["ms"], 100 * r["ms"] / tot,
            "%.1f" % tf if tf is not None else "-",
            "%.1f%%" % (100 * r["frac_of_bf16_peak"]) if r.get("frac_of_bf16_peak") is not None else "-",
            "%.0f" % r["gbs"] if r.get("gbs") is not None else "-",
            "%.1f%%" % (100 * r["frac_of_hbm_peak"]) if r.get("frac_of_hbm_peak") is not None else "-"))
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "r01")
