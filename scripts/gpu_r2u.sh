#!/bin/bash
# Round 2, GPU call U (1 GPU): validation of the default configuration (read2sdbg, item pruning, folded mercy filter, fused
# build in rounds) + A/B of the folded filter + wide k + ncu evidence for the new kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rxXf > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2u_pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2u_bench_default.json 2> gpurun_out/r2u_bench_default.err; tail -2 gpurun_out/r2u_bench_default.err
echo "== bench without the folded mercy filter"
MHB_MERCY_FOLD=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2u_bench_nofold.json 2> gpurun_out/r2u_bench_nofold.err; tail -2 gpurun_out/r2u_bench_nofold.err
python - <<PY
import json
for f in ("default", "nofold"):
    try:
        j = json.loads([l for l in open('gpurun_out/r2u_bench_%s.json' % f) if l.startswith('{')][-1]); r = j['roofline']
        print(f, 'ms/step %.1f value %.3g e2e %.3g (%.1f ms) pass %.2f ms frac %.3f launches %s' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], r['avg_launch_ms'], r['frac'], j['gpu_launches']))
        print('  per pass frac', [round(x, 3) for x in r['per_pass_frac']], 's2s', round(r['s2s_pass']['frac'], 3), j['config'].get('host_affinity'), 'clocks', j.get('clocks'))
        print('  ', {k: round(v, 1) for k, v in j['stage_ms'].items()}, {k: round(v, 1) for k, v in j['e2e']['stages'].items() if isinstance(v, float)})
    except Exception as e:
        print(f, 'unreadable', e)
PY
echo "== bench reference arm"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
for k in 99 119 141; do
  timeout 300 python bench.py --k $k --reads 5000000 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/klist_k$k.json 2> gpurun_out/klist_k$k.err
  python - $k <<PY
import json, sys
k = sys.argv[1]
try:
    j = json.loads([l for l in open('gpurun_out/klist_k%s.json' % k) if l.startswith('{')][-1])
    r = j['roofline']
    print('k=%s: %.1f ms/step  %.3g edges/s  e2e %.3g  records %d B, %d passes, pass %.2f ms frac %.3f  s2s items %d pass frac %.3f  stages %s' % (
        k, j['ms_per_step'], j['value'], j['e2e']['value'] or 0, r['algorithmic_bytes_per_launch'] // 2 // j['config']['n_edge_records'],
        len(r['per_pass_ms']), r['avg_launch_ms'], r['frac'], j['config']['n_sdbg_sort_items'], r['s2s_pass']['frac'], {a: round(b, 1) for a, b in j['stage_ms'].items()}))
except Exception as e:
    print('k=%s unreadable' % k, e, open('gpurun_out/klist_k%s.err' % k).read()[-300:])
PY
done
echo "== ncu launch list at bench size"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2u_launch_list_10Mreads.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu --set full: pruned seq2sdbg extraction, folded mercy marks, stable radix pass at bench size"
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_s2s_extract_edges_pruned|k_mark_mercy_roll_fold|k_radix_pass3' -c 4 -o gpurun_out/r2u_new_kernels_10M \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/r2u_new_kernels_10M.ncu-rep 2>/dev/null
