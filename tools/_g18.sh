mkdir -p gpurun_out/r2
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/r2/bench_final_n8.json 2> gpurun_out/r2/bench_final_n8.err; echo "n8 rc=$?"; tail -4 gpurun_out/r2/bench_final_n8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2/bench_final_n8.json").read().strip().splitlines()[-1])
print("N=8 headline ms/step %.2f value %.4g e2e %.4g vec %.4g frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e_per_replica_vectors"]["value"], d["roofline"]["frac"]), d["clocks"])
for e in d["extra_configs"]:
    if "error" in e: print(e); continue
    print(e["key"], "ms/step %.2f value %.4g e2e %.4g frac %.3f" % (e["ms_per_step"], e["value"], e["e2e"]["value"], e["roofline"]["frac"]), e["replicas_total"])
PY
