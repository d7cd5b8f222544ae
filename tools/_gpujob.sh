export VB_LIB_PATH=$PWD/vslam_b200/lib_tuning/libvslam_b200.so
TC_DRAINS=6,8 timeout 300 python tools/tc_variants.py 2>&1 | tail -2 | tee gpurun_out/r2as_tc_variants.log
TC_DRAINS=6,8 TC_EXTRA=tc_svc_hi=1 timeout 300 python tools/tc_variants.py 2>&1 | tail -2 | tee -a gpurun_out/r2as_tc_variants.log
