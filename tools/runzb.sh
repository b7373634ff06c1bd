for zb in 512 1024; do
  echo -n "zb=$zb "
  PF_KS_BATCH=$zb timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/zb_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],3), d['phases_ms_per_step'])" || tail -3 gpurun_out/zb_err.log
done
