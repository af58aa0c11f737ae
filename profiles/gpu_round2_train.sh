#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_block.py tests/test_abi.py -m gpu -x -q > gpurun_out/train_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/train_pytest.log
tail -15 gpurun_out/train_pytest.log
timeout 600 python profiles/bench_training_block.py > gpurun_out/training_block.json 2> gpurun_out/training_block.err
cat gpurun_out/training_block.json; tail -3 gpurun_out/training_block.err
