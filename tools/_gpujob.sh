for sv in 0 1; do TC_DRAINS=0,1,4 TC_EXTRA=tc_svc_hi=$sv python tools/tc_variants.py; done > gpurun_out/r2m_tc_svc.log 2>&1; cat gpurun_out/r2m_tc_svc.log
