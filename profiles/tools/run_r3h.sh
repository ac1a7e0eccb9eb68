set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SPLISER_TIMING=1 timeout 600 python profiles/tools/bam_ingest_profile.py 8000000 seq > gpurun_out/r3h_seq.log 2>&1
tail -12 gpurun_out/r3h_seq.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3h_launches_seq.csv python profiles/tools/bam_ingest_profile.py 8000000 seq > gpurun_out/r3h_ncu.log 2>&1
SPLISER_TIMING=1 timeout 600 python profiles/tools/bam_ingest_profile.py 40000000 > gpurun_out/r3h_plain.log 2>&1
tail -6 gpurun_out/r3h_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3h_launches_plain.csv python profiles/tools/bam_ingest_profile.py 40000000 > gpurun_out/r3h_ncu2.log 2>&1
