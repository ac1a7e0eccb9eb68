# final artifacts of the build with the cooperative K1: full GPU suite, smoke, bench line, reference arm, ncu launch list, ncu --set full of every kernel of the step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -8 > gpurun_out/f2_pytest.log
cat gpurun_out/f2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err
tail -2 gpurun_out/f2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f2_reference.json 2> gpurun_out/f2_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f2_launches_c2.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/f2_ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gb_coop|k_gb_sb_fill|k_chunk_bounds|k_count_fused|k_hot_items|k_span_blocksum|k_finalize" -s 28 -c 7 -f -o gpurun_out/f2_step python bench.py --profile --steps 2 --warmup 3 > gpurun_out/f2_ncu_step.log 2>&1
ls -la gpurun_out | tail -8
