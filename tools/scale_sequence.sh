set -x
for N in 1 2 4 8; do
  P=$((29600+N))
  if [ $N = 1 ]; then
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
    python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+50)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_n$N.json")); r=json.load(open("gpurun_out/scale_ref_n$N.json"))
print("N=$N", round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), "pcm16", round(d["e2e_pcm16"]["value"]), "ragged", round(d["e2e_ragged"]["value"]), "strong", d.get("strong",{}).get("efficiency"), "ref", round(r["value"]), r.get("impl"))
PY
done
