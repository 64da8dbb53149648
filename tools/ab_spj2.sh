mkdir -p gpurun_out
for spj in 0 2 3 5; do
  python bench.py --workload C2 --steps 30 --warmup 5 --no-cpu-baseline --no-extras --sg-per-job $spj > gpurun_out/ab_spj$spj.json 2> gpurun_out/ab_spj$spj.err
  echo "C2 spj=$spj rc=$? $(python -c "import json;d=json.load(open('gpurun_out/ab_spj$spj.json'));print(round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']),round(d['roofline_frac'],3))" 2>&1 | tail -1)"
done
for spj in 0 8 32; do
  python bench.py --workload C3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --sg-per-job $spj > gpurun_out/ab_spj3_$spj.json 2> gpurun_out/ab_spj3_$spj.err
  echo "C3 spj=$spj rc=$? $(python -c "import json;d=json.load(open('gpurun_out/ab_spj3_$spj.json'));print(round(d['value']),round(d['ms_per_step'],3),round(d['e2e']['value']),round(d['roofline_frac'],3))" 2>&1 | tail -1)"
done
