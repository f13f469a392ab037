ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches52.csv python tools/profile_frame.py > gpurun_out/ncu52a.log 2>&1
python tools/launch_summary.py gpurun_out/launches52.csv > gpurun_out/launches52_summary.txt
cat gpurun_out/launches52_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_gemm_s3 -s 6 -c 1 -o gpurun_out/s3chain_dcb_v5 -f python tools/one_dcb.py 160 240 256 > gpurun_out/ncu52b.log 2>&1
tail -2 gpurun_out/ncu52b.log
python tools/gemm_probe.py dw dcb > gpurun_out/probe52.log 2>&1; cat gpurun_out/probe52.log
