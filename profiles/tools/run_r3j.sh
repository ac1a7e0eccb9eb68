set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/r3j_pytest.log
cat gpurun_out/r3j_pytest.log
timeout 600 python profiles/tools/bam_ingest_profile.py 8000000 seq 2>&1 | tail -2
timeout 600 python profiles/tools/bam_ingest_profile.py 40000000 2>&1 | tail -2
timeout 300 python profiles/tools/quick_time.py c2 40000000 fused 2>/dev/null
