#!/bin/bash
# GPU session: evidence for the fused last layer -- sustained A/B, bench with / without the steady-state pre-roll,
# compute-sanitizer on the tiny config, ncu launch list + full captures (wide layer, fused last layer), SASS histograms
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s13}
echo "== sustained probe (full-sector c loads)"
timeout 300 python tools/power_probe.py --seconds 3 --what "enc,enc" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-700 $O/power_$TAG.jsonl
echo "== bench, no pre-roll / pre-roll (short form)"
BENCH_PREROLL_S=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/bench_nopre_$TAG.json 2> $O/bench_nopre_$TAG.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/bench_pre_$TAG.json 2> $O/bench_pre_$TAG.err; echo "rc=$?"
echo "== bench (driver form)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
for f in ('bench_nopre','bench_pre','bench'):
    try:
        d=json.loads(open('$O/%s_$TAG.json' % f).read().strip().splitlines()[-1])
        print(f, 'value %.0f  e2e %.0f  single %.0f  ms/step %.2f  frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['single_batch']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac']), d['clocks'])
        print('   phases', d['roofline']['phase_ms_last_call'], d['roofline']['phase_sm_mhz'])
    except Exception as e: print(f, 'no line', e)
PY
export IE_SPIN_LIMIT_MS=600000
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_tiny.py > $O/sanitizer_${tool}_$TAG.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_tiny ok|rel_l2" $O/sanitizer_${tool}_$TAG.log | tail -7
done
unset IE_SPIN_LIMIT_MS
echo "== ncu launch list (one 1280 x 512 encode)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_list_$TAG.log 2>&1
echo "rc=$?"; tail -1 $O/ncu_list_$TAG.log | cut -c1-300
python - <<PY
import csv
rows=[r for r in csv.reader(open('$O/launches_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki].split('(')[0][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
tot=sum(v for _,v in agg.values())
for k,(n,v) in sorted(agg.items(), key=lambda x:-x[1][1]): print('%-62s n=%4d  %10.3f ms  %5.1f %%' % (k,n,v/1e6,100*v/tot))
PY
echo "== ncu --set full: wide layer (layer 1) and the fused last layer"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_kernel -s 5 -c 1 -o $O/prof_layer_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_layer_$TAG.log 2>&1
echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_fused_kernel -s 1 -c 1 -o $O/prof_fused_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_fused_$TAG.log 2>&1
echo "rc=$?"
for k in layer fused; do
  ncu -i $O/prof_${k}_$TAG.ncu-rep --page raw --csv > $O/ncu_full_${k}_$TAG.csv 2>/dev/null
  python - <<PY
import csv
try:
    rows=list(csv.reader(open('$O/ncu_full_${k}_$TAG.csv')))
    d=dict(zip(rows[0],rows[2]))
    for key in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum','gpc__cycles_elapsed.max.per_second'):
        print('$k', key, d.get(key,'')[:90])
    for key,v in d.items():
        if key.endswith('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'): print('$k', key, v)
except Exception as e: print('$k no capture', e)
PY
done
echo "== SASS histograms"
cuobjdump -sass code_intelligence_b200/libissue_emb_b200.so > $O/sass_$TAG.txt 2>/dev/null
python - <<PY
import re,collections,json
txt=open('$O/sass_$TAG.txt').read()
out={}
for m in re.finditer(r'Function : (\S+)(.*?)(?=Function : |\Z)', txt, re.S):
    name=m.group(1); body=m.group(2)
    ops=collections.Counter(re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', body, re.M))
    out[name]={k:v for k,v in ops.items() if any(s in k for s in ('UTCHMMA','UTMALDG','UTMASTG','LDTM','UTCBAR','SYNCS','MUFU','RED','LDG','STG','STS','UBLKCP','ATOM','NANOSLEEP','BAR','HMMA'))}
json.dump(out, open('$O/sass_hist_$TAG.json','w'), indent=1)
for n,h in out.items():
    if 'fused' in n: print(n[:80], {k:v for k,v in sorted(h.items()) if k.split('.')[0] in ('UTCHMMA','UTMALDG','UTMASTG','LDTM','MUFU','RED','UTCBAR','LDG')})
PY
rm -f $O/sass_$TAG.txt
