timeout 300 python tools/elastic_gpu_check.py resnet50 2 2 > gpurun_out/elastic_n2_warm.log 2>&1; echo "elastic warm rc=$?"; grep "^rescale" gpurun_out/elastic_n2_warm.log | cut -c1-500
grep -h "joiner ready\|rescaled to" gpurun_out/elastic_logs_resnet50_n2_pool2/*.log | head
timeout 300 python tools/elastic_gpu_check.py resnet50 2 2 NCCL_NVLS_ENABLE=0 > gpurun_out/elastic_n2_warm_nonvls.log 2>&1; echo "elastic warm nonvls rc=$?"; grep "^rescale" gpurun_out/elastic_n2_warm_nonvls.log | cut -c1-500
grep -h "joiner ready\|rescaled to" gpurun_out/elastic_logs_resnet50_n2_pool2/*.log | head
