# Round profile on the GPU box:  bash tools/profile_round.sh <tag>
#   launch list of one steady-state `performance` P frame at 1920x1280 (real operands: random-init weights, synthetic clip),
#   one `ncu --set full` capture of a DepthConvBlock-256 chain launch INSIDE that frame (third chain launch of the frame:
#   dc.3 -> ffn.0 -> ffn.2 -> next dc.0), one capture each of the HBM-bound kernels, single-layer probes.
tag=${1:-rXX}
export DMC_GRAPH=0        # (kernel nodes of a graph launch are profiled too, but the skip counts below are simpler without)
python tools/profile_frame.py > gpurun_out/frame_$tag.log 2>&1 && tail -1 gpurun_out/frame_$tag.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_frame.py > gpurun_out/ncu_launch_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launches_${tag}_summary.txt
cat gpurun_out/launches_${tag}_summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_s3 -s 2 -c 1 -o gpurun_out/chain_dcb_$tag -f python tools/profile_frame.py > gpurun_out/ncu_chain_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_chain_$tag.log
ncu -i gpurun_out/chain_dcb_$tag.ncu-rep --page raw --csv > gpurun_out/chain_dcb_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/chain_dcb_$tag.ncu-rep --page source --csv > gpurun_out/chain_dcb_${tag}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/chain_dcb_${tag}_raw.csv
python tools/stall_summary.py gpurun_out/chain_dcb_${tag}_src.csv 20 > gpurun_out/chain_dcb_${tag}_stall.txt 2>&1; head -12 gpurun_out/chain_dcb_${tag}_stall.txt
ncu --set full --clock-control none --profile-from-start off -k "regex:k_prior_step|k_prior_finish|k_dwconv3x3_strip|k_frame_stats|k_round_z_bits|k_frames_from_u8" -c 40 -o gpurun_out/hbm_kernels_$tag -f python tools/profile_frame.py > gpurun_out/ncu_hbm_$tag.log 2>&1
ncu -i gpurun_out/hbm_kernels_$tag.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/hbm_kernels_${tag}_raw.csv
python tools/gemm_probe.py 2 dw dcb > gpurun_out/probe_$tag.log 2>&1; cat gpurun_out/probe_$tag.log
