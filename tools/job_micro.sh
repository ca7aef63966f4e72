./tools/micro/mma_rate > gpurun_out/r2m_mma_rate.txt 2>&1
cat gpurun_out/r2m_mma_rate.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ssim" -c 2 -o gpurun_out/r2m_ssim -f python bench.py --steps 1 --warmup 1 --no-comparators > gpurun_out/r2m_ncu.log 2>&1
tail -2 gpurun_out/r2m_ncu.log
ls -la gpurun_out/r2m_ssim.ncu-rep
