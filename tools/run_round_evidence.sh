# Round evidence on one B200: bench line, ncu launch list, full ncu capture of one model step, sanitizers.
#   bash tools/run_round_evidence.sh <tag>        (outputs under gpurun_out/<tag>_*)
set -x
TAG=${1:-r2k}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_nbr_search|k_node_encode_tc|k_edge_encode_tmem|k_edge_agg|k_node_update_tc" -c 9 -f -o gpurun_out/${TAG}_full python tools/ncu_probe.py 1024 300 1 > gpurun_out/${TAG}_full.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; tail -n 3 gpurun_out/${TAG}_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/${TAG}_sanitizer_synccheck.txt 2>&1; tail -n 3 gpurun_out/${TAG}_sanitizer_synccheck.txt
ls -la gpurun_out | tail -8
