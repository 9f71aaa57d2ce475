mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r2_t1.log
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
cat gpurun_out/r2_bench1.json | cut -c1-600
