#!/bin/bash
# r02m: launch lists (B=64, B=8) of the final state, tensor-pipe counters + DRAM traffic of the GEMM kernels of one decode step
set -u
TAG=r02m
mkdir -p gpurun_out
Q="--no-cpu --no-parity --eager-gpu 0"
KPS=1592
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 2)) -c $KPS --csv \
    --log-file gpurun_out/${TAG}_b64_launches.csv python bench.py --steps 1 --warmup 3 $Q > gpurun_out/${TAG}_b64_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_b64_launches.csv gpurun_out/${TAG}_b64_launches_summary.csv | head -30
KPS8=1430
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS8 * 2)) -c $KPS8 --csv \
    --log-file gpurun_out/${TAG}_b8_launches.csv python bench.py --batch 8 --steps 1 --warmup 3 $Q > gpurun_out/${TAG}_b8_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_b8_launches.csv gpurun_out/${TAG}_b8_launches_summary.csv | head -30
# tensor-pipe counters for the tensor-core kernels of one decode step (around step 16): explicit metric list, not --set full
M="gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_2cta.sum,sm__inst_executed_pipe_tmem.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum"
timeout 900 ncu --metrics $M --clock-control none -k regex:"tc_gemm_kernel" -s 700 -c 26 --csv \
    --log-file gpurun_out/${TAG}_tensor_counters.csv python bench.py --steps 1 --warmup 3 $Q > gpurun_out/${TAG}_tc_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02m_tensor_counters.csv", errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[0]; iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
by = collections.OrderedDict()
for r in rows[1:]:
    by.setdefault(r[iid], {"kernel": r[ik][:60]})[r[im]] = (r[iv], r[iu])
import json
out = []
for k, d in by.items():
    g = lambda n: float(d[n][0].replace(",", "")) if n in d and d[n][0] not in ("", "n/a") else float("nan")
    out.append({"id": k, "kernel": d["kernel"], "time": d.get("gpu__time_duration.sum"), "cycles": g("sm__cycles_elapsed.avg"),
                "hmma_inst": g("sm__inst_executed_pipe_tensor_subpipe_hmma.sum"),
                "hmma_active_pct": g("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                "tensor_active_pct": g("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                "tmem_active_pct": g("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "dram_read": d.get("dram__bytes_read.sum"), "dram_write": d.get("dram__bytes_write.sum")})
json.dump(out, open("gpurun_out/r02m_tensor_counters.json", "w"), indent=1)
for o in out[:26]:
    print(o)
PY
