"""profiles/r2_traffic.json from an ncu --set full report of tools/quick_stage0.py 8 (one launch = 8 frames):
python tools/ncu_traffic.py gpurun_out/r2_full.ncu-rep profiles/r2_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, key, scale_units=True):
    if key not in ix or r[ix[key]] in ("", "n/a"):
        return None
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v


names, stage_i = {}, 0
res = {"source": f"ncu --set full --clock-control none, tools/quick_stage0.py 8 (one launch = 8 frames of 1920x1080), round-2 kernels; summary: profiles/r2_ncu_full.txt; launch list of `bench.py --steps 2 --warmup 3`: profiles/r2_launches.csv",
       "frames_per_launch": 8, "dram_bytes_per_launch": {}, "l2_to_l1_bytes_per_launch": {}, "ncu_pct_of_peak": {}, "global_load_requests_per_launch": {}, "wavefronts_per_global_load_request": {}}
for r in rows[2:]:
    k = r[ix["Kernel Name"]]
    if "k_scan_stage0<" in k: name = "k_scan_stage0_even"
    elif "k_scan_odd" in k: name = "k_scan_odd"
    elif "k_integral_walk" in k: name = "k_integral_walk"
    elif "k_strip_carry" in k: name = "k_strip_carry"
    elif "k_cell_bounds" in k: name = "k_cell_bounds"
    elif "k_scan_stage<" in k:
        name = ["k_scan_stage_exact0", "k_scan_stage_1", "k_scan_stage_2", "k_scan_stage_3"][min(stage_i, 3)]; stage_i += 1
    elif "k_replay_rows" in k: name = "k_replay_rows"
    elif "k_row_events" in k: name = "k_row_events"
    elif "k_finalize" in k: name = "k_finalize"
    else: continue
    if name in res["dram_bytes_per_launch"]:
        continue
    res["dram_bytes_per_launch"][name] = int((num(r, "dram__bytes_read.sum") or 0) + (num(r, "dram__bytes_write.sum") or 0))
    x = num(r, "l1tex__m_xbar2l1tex_read_bytes.sum")
    if x: res["l2_to_l1_bytes_per_launch"][name] = int(x)
    req = num(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", False)
    wav = num(r, "l1tex__data_pipe_lsu_wavefronts_mem_global_op_ld.sum", False) or num(r, "l1tex__data_pipe_lsu_wavefronts.sum", False)
    if req: res["global_load_requests_per_launch"][name] = int(req)
    if req and wav: res["wavefronts_per_global_load_request"][name] = round(wav / req, 2)
    res["ncu_pct_of_peak"][name] = {
        "l1tex_data_pipe": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
        "lts_throughput": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "lts_t_sectors_pct": num(r, "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", False),
        "issue_active": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "dram_throughput": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "l1_hit_rate": num(r, "l1tex__t_sector_hit_rate.pct", False), "l2_hit_rate": num(r, "lts__t_sector_hit_rate.pct", False),
        "ms": num(r, "gpu__time_duration.sum"), "registers": num(r, "launch__registers_per_thread", False),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False)}
res["ncu_pct_of_peak"]["k_scan_stage0_even"]["lts_t_sectors_note"] = "lts__t_sectors.sum.pct_of_peak_sustained_elapsed: the L2 sector-throughput figure BASELINE.json's 60 % target names; lts__throughput (busiest L2 sub-unit) is lts_throughput"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["ncu_pct_of_peak"]["k_scan_stage0_even"]), res["dram_bytes_per_launch"], res["wavefronts_per_global_load_request"])
