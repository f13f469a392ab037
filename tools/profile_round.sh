# Round profile: launch list of one steady-state P frame, one `ncu --set full` capture of the DCB-256 chain launch,
# single-layer probes.  Run on the GPU box:  bash tools/profile_round.sh <tag>
tag=${1:-rXX}
python tools/profile_frame.py > gpurun_out/frame_$tag.log 2>&1 && cat gpurun_out/frame_$tag.log | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_frame.py > gpurun_out/ncu_launch_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launches_${tag}_summary.txt
cat gpurun_out/launches_${tag}_summary.txt
python tools/one_dcb.py 160 240 256 > gpurun_out/one_dcb_$tag.log 2>&1 && tail -1 gpurun_out/one_dcb_$tag.log
ncu --set full --clock-control none --import-source on -k regex:k_gemm_s3 -s 6 -c 1 -o gpurun_out/chain_dcb_$tag -f python tools/one_dcb.py 160 240 256 > gpurun_out/ncu_chain_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_chain_$tag.log
python tools/gemm_probe.py 2 dw dcb > gpurun_out/probe_$tag.log 2>&1; cat gpurun_out/probe_$tag.log
