set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01b_bench_ref.json 2>/dev/null; echo "ref rc=$?"
ncu --kernel-name regex:"fuse|radix|bracket|cand|apply" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 40 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:"fuse_sources_tma|bracket_classify" --launch-skip 2 --launch-count 2 -o gpurun_out/r01b_k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images-per-gpu 2000 > gpurun_out/ncu_f.log 2>&1; echo "ncu full(2000) rc=$?"
if [ ! -s gpurun_out/r01b_k1_full.ncu-rep ]; then
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:"fuse_sources_tma|bracket_classify" --launch-skip 2 --launch-count 2 -o gpurun_out/r01b_k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images-per-gpu 200 > gpurun_out/ncu_f.log 2>&1; echo "ncu full(200) rc=$?"
fi
ls -la gpurun_out/*.ncu-rep
