# Round evidence at N=1 (run under gpurun): tests, smoke, bench line, reference arm, ncu launch list of the bench command,
# ncu --set full of K1 + classify pass at full size and of the loss / metric kernels.  Outputs under gpurun_out/ (prefix $P).
P=${1:-r01c}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_ref.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/${P}_bench_n1.json 2> gpurun_out/${P}_bench_n1.err; echo "bench rc=$?"
python tools/bench_extra.py > gpurun_out/${P}_bench_extra.jsonl 2> gpurun_out/${P}_bench_extra.err; echo "extra rc=$?"
ncu --kernel-name regex:"fuse|radix|bracket|cand|apply" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 40 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:"uw_ce_fused|miou_kernel" --launch-skip 4 --launch-count 8 -o gpurun_out/${P}_k4_full -f python tools/bench_extra.py --section loss --quick > gpurun_out/ncu_k4.log 2>&1; echo "ncu full(K4) rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name regex:"fuse_sources_tma|bracket_classify" --launch-skip 2 --launch-count 2 -o gpurun_out/${P}_k1_full -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --images-per-gpu 2000 > gpurun_out/ncu_f.log 2>&1; echo "ncu full(2000) rc=$?"
ls -la gpurun_out/*.ncu-rep
