cd $GRAFT_REPO_ROOT
for i in 1 2; do timeout 100 python -m pytest tests/test_gpu_step.py -q -m gpu -k "c_fused_esat_step_bf16" 2>&1 | grep -E "assert|Error|passed|failed|hist|^E " | head -12; done
ADVMIL_ESAT_OVERLAP=0 timeout 100 python -m pytest tests/test_gpu_step.py -q -m gpu -k "c_fused_esat_step_bf16" 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -8
