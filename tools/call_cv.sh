#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "costvolume or golden or ranges or midsize or stereo or cli" > gpurun_out/cv_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/cv_pytest.log
tail -6 gpurun_out/cv_pytest.log
python tools/micro/cv_time.py 2>&1 | tail -5
MGMB200_CV_WARP_PER_PIXEL=1 python tools/micro/cv_time.py 2>&1 | tail -5
