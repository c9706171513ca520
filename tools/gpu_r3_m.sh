#!/bin/bash
# Round 2, third session, GPU call M (1 GPU): the whole GPU suite with the pair kernel's activation warps switched ON for every context (env AGRS_TC_ACT_WARPS=1)
mkdir -p gpurun_out
AGRS_TC_ACT_WARPS=1 timeout 150 python -m pytest tests -m gpu -q -x -k "not default_is_the_second_pass" > gpurun_out/r3m_pytest_act_warps_on.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3m_pytest_act_warps_on.log
