# Round evidence at N=1 (run under gpurun): tests, smoke, bench line, reference arm, ncu launch list of the bench command.
# Outputs under gpurun_out/ (prefix $P).
P=${1:-r02}
set -x
python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${P}_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err; echo "bench rc=$?"
ncu --kernel-name regex:"fuse|radix|bracket|cand|apply" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 40 --csv --log-file gpurun_out/${P}_bench_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
