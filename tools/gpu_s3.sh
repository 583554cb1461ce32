#!/bin/bash
# GPU session: compute-sanitizer on the tiny config, ncu launch list + full captures, SASS histograms, seq_len sweep
set -u
mkdir -p gpurun_out
O=gpurun_out; TAG=${1:-s3}
export IE_SPIN_LIMIT_MS=600000
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_tiny.py > $O/sanitizer_${tool}_$TAG.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_tiny ok|rel_l2" $O/sanitizer_${tool}_$TAG.log | tail -6
done
unset IE_SPIN_LIMIT_MS
echo "== ncu launch list (one 1280 x 512 encode)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_list_$TAG.log 2>&1
echo "rc=$?"; tail -2 $O/ncu_list_$TAG.log | cut -c1-300
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
echo "== ncu --set full: lstm_layer_kernel (layer 1 of a 1280 x 512 encode) and the pair GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_kernel -s 5 -c 1 -o $O/prof_layer_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_layer_$TAG.log 2>&1
echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair -s 4 -c 1 -o $O/prof_gemm_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_gemm_$TAG.log 2>&1
echo "rc=$?"
for k in layer gemm; do
  ncu -i $O/prof_${k}_$TAG.ncu-rep --page raw --csv > $O/ncu_full_${k}_$TAG.csv 2>/dev/null
  python - <<PY
import csv
rows=list(csv.reader(open('$O/ncu_full_${k}_$TAG.csv')))
if len(rows)>=3:
    h,u,v=rows[0],rows[1],rows[2]
    want=['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','derived__lts__lts2xbar_bytes.sum.per_second','gpc__cycles_elapsed.max.per_second','sm__warps_active.avg.per_cycle_active','sm__inst_executed.avg.per_cycle_elapsed','launch__registers_per_thread']
    for a,b,c in zip(h,u,v):
        if any(w==a or (w in a and 'TriageCompute' not in a and len(a)<len(w)+3) for w in want): print('$k',a,b,c[:80])
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
    key=re.sub(r'^_ZN\d*','',name)[:70]
    out[name]={k:v for k,v in ops.items() if any(s in k for s in ('UTCHMMA','UTMALDG','UTMASTG','LDTM','UTCBAR','SYNCS','MUFU','RED','LDG','STG','STS','UBLKCP','ATOM','NANOSLEEP','BAR','HMMA'))}
json.dump(out, open('$O/sass_hist_$TAG.json','w'), indent=1)
for n,h in out.items():
    if 'lstm' in n or 'gemm_bf16_pair' in n: print(n[:80], {k:v for k,v in sorted(h.items()) if k.split('.')[0] in ('UTCHMMA','UTMALDG','UTMASTG','LDTM','MUFU','RED','UTCBAR')})
PY
echo "== seq_len sweep (configs[2])"
timeout 900 python tools/sweep_seq_len.py > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err; echo "rc=$?"; cat $O/sweep_$TAG.jsonl | cut -c1-400
