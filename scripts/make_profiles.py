"""Turn the ncu outputs a GPU trip left under gpurun_out/ncu/ (scripts/ncu_step.sh, scripts/ncu_gemm.sh) into the
tracked summaries under profiles/: the per-kernel launch list of one train step (+ DRAM traffic per GEMM launch, which
bench.py reports as roofline.traffic) and the --set full summary of the dominant kernel."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCU = os.path.join(ROOT, "gpurun_out", "ncu")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_tc_v3"


def launches():
    rows = [r for r in csv.reader(open(os.path.join(NCU, "launches.csv"))) if len(r) > 10]
    ci = {h: i for i, h in enumerate(rows[0])}
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]]})
        v, u, m = float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]], r[ci["Metric Name"]]
        if m.startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        else:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
        d[m] = v
    L = list(L.values())
    idx = [i for i, d in enumerate(L) if "tc_prep_input" in d["name"]]
    step = L[idx[1]:idx[2]] if len(idx) > 2 else L[idx[-1]:]

    def short(n):
        m = re.search(r"(tc_gemm_kernel<[^>]*>|[a-z_0-9]+_kernel(<[^>]*>)?)", n)
        return (m.group(1) if m else n[:40]).replace("(bool)", "").replace("(int)", "")
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(short(d["name"]), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d["gpu__time_duration.sum"]
        a["rd"] += d["dram__bytes_read.sum"]
        a["wr"] += d["dram__bytes_write.sum"]
    tot = sum(a["us"] for a in agg.values())
    out = [f"# Round 1, tensor-core engine ({tag}: persistent tcgen05 3xTF32 segment-GEMM, cta_group::2 pairs for forward, dgrad and wgrad) — ncu launch list, one train step", "",
           "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv python scripts/one_step.py --steps 3` (scripts/ncu_step.sh)",
           "(C2: 4096 patches 7x7x145, 15 classes; cold-cache, serialised: compare shares, not absolutes).  One step = the launches between two `tc_prep_input_kernel`.",
           "`tc_gemm_kernel<MN, CG, EPI>`: MN=0 K-major (forward EPI=0 store + BN statistics, dgrad EPI=1 accumulate), MN=1 MN-major wgrad (EPI=2 atomic); CG = CTA group size.", "",
           "| kernel | launches | total us | share | DRAM read MB | DRAM write MB | avg DRAM GB/s |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"| `{k}` | {a['n']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {(a['rd'] + a['wr']) / a['us'] / 1e3:.0f} |")
    gem = [d for d in step if "tc_gemm_kernel" in d["name"]]
    gt = sum(d["gpu__time_duration.sum"] for d in gem)
    gb = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in gem)
    out += ["", f"Total {tot / 1e3:.2f} ms over {len(step)} launches.",
            f"GEMM launches: {len(gem)}, {gt / 1e3:.2f} ms ({100 * gt / tot:.1f} % of the step), DRAM traffic {gb / 1e9:.2f} GB per step = {gb / len(gem) / 1e6:.1f} MB per launch (average)."]
    open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md"), "w").write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:   # the raw per-launch list of that step
        f.write("launch,kernel,gpu_time_us,dram_read_bytes,dram_write_bytes\n")
        for i, d in enumerate(step):
            f.write(f"{i},\"{short(d['name'])}\",{d['gpu__time_duration.sum']:.2f},{d['dram__bytes_read.sum']:.0f},{d['dram__bytes_write.sum']:.0f}\n")
    json.dump({"gemm_launches_per_step": len(gem), "gemm_dram_bytes_per_step": gb, "gemm_dram_bytes_per_launch": gb / len(gem),
               "gemm_share_of_step_ncu": gt / tot, "source": f"profiles/{tag}_launches_summary.md"},
              open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json"), "w"), indent=1)
    print("\n".join(out[-3:]))


def full():
    names = ["fwd_conv_dec_0", "fwd_connector_1", "wgrad_connector_1"]
    keys = [("gpu__time_duration.sum",) * 2, ("dram__bytes_read.sum",) * 2, ("dram__bytes_write.sum",) * 2,
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",) * 2, ("dram__bytes_read.sum.per_second",) * 2,
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed",) * 2, ("sm__throughput.avg.pct_of_peak_sustained_elapsed",) * 2,
            ("sm__pipe_tensor_cycles_active_realtime (pct of peak, TPC)", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum",) * 2, ("l1tex__m_l1tex2xbar_write_bytes.sum",) * 2,
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",) * 2, ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",) * 2,
            ("launch__registers_per_thread",) * 2, ("launch__shared_mem_per_block_dynamic",) * 2, ("launch__grid_size",) * 2,
            ("launch__cluster_size",) * 2, ("smsp__cycles_active.avg",) * 2, ("smsp__inst_executed.sum",) * 2,
            ("lts__t_sector_hit_rate.pct",) * 2]
    vals = {}
    for f in names:
        rows = list(csv.reader(open(os.path.join(NCU, f + ".raw.csv"))))
        vals[f] = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    kern = {f: re.sub(r"\(bool\)|\(int\)", "", vals[f].get("Kernel Name", ("", ""))[0]) for f in names}
    out = ["# Round 1 — `ncu --set full` of the dominant kernel, `tc_gemm_kernel` (persistent tcgen05 3xTF32 segment-GEMM)", "",
           "Command (per launch): `ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s <59 + idx> -c 1 python scripts/one_step.py --steps 2` (scripts/ncu_gemm.sh);",
           "C2 shape, batch 4096.  Three launches of the second train step: the largest 1x1 conv forward, the 240->4x30 level forward, and its wgrad.", "",
           "| metric | " + " | ".join(f"{f}" for f in names) + " |", "|---|---|---|---|"]
    for label, key in keys:
        r = []
        for f in names:
            v, u = vals[f].get(key, ("n/a", ""))
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            r.append(f"{v} {u}".strip())
        out.append(f"| `{label}` | " + " | ".join(r) + " |")
    out += ["", "Reading: `sm__pipe_tensor_cycles_active_realtime` is normalised to the 16-bit-input rate, so a kind::tf32 kernel saturates the pipe at 50 %;",
            "`l1tex__m_xbar2l1tex_read_bytes` is the L2 -> SM operand traffic (TMA), several times the DRAM bytes: the operands are re-read from L2 by design (taps, N tiles).",
            "SASS evidence (cuobjdump -sass hypelcnn_b200/lib/libhypelcnn_b200.so): `UTCHMMA.2CTA` (tcgen05.mma cta_group::2), `UTMALDG.4D.2CTA` (TMA), `LDTM.x32` (tcgen05.ld), `UTCBAR.2CTA.MULTICAST` (tcgen05.commit multicast), `UTCATOMSWS.2CTA` (TMEM alloc).",
            f"In-kernel role timing of every GEMM launch of a step (HYP_TC_TIMING=1): `profiles/{tag}_role_timing.txt`."]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w").write("\n".join(out) + "\n")


launches()
full()
src = os.path.join(ROOT, "gpurun_out", "tc_timing.txt")
if os.path.exists(src):
    open(os.path.join(ROOT, "profiles", f"{tag}_role_timing.txt"), "w").write(open(src).read())
bl = os.path.join(ROOT, "gpurun_out", "bench.log")
if os.path.exists(bl):
    open(os.path.join(ROOT, "profiles", f"{tag}_bench.json"), "w").write(open(bl).read().strip().split("\n")[-1] + "\n")
