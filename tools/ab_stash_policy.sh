for v in p0 p1 p2 p3; do
  export PDE_B200_LIB=$PWD/neural-network-based-pde-solver_b200/libpde_b200_$v.so
  echo "== $v"; python tools/prof_step.py 22 5
  python -m pytest tests/test_gpu_tc.py -m gpu -q -p no:cacheprovider -k "large_batch or full_size or adjoint_scale" 2>&1 | tail -1
done
for v in p0 p1 p2 p3; do
  export PDE_B200_LIB=$PWD/neural-network-based-pde-solver_b200/libpde_b200_$v.so
  ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:tc_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_dram_$v.csv python tools/prof_step.py 22 3 > /dev/null 2>&1
  echo "== $v"; grep -E "dram__bytes|gpu__time|lts__t" gpurun_out/r2_dram_$v.csv | awk -F, '{print $(NF-2), $(NF-1), $NF}'
done
