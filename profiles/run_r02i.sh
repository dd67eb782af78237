set -x
timeout 300 python -m pytest tests/test_mf_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "planned or model or fit" 2>&1 | tail -4
bash profiles/run_variants.sh base v5 v4 vpl1
