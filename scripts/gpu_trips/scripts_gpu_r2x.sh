#!/bin/bash
# Round 2, two GPUs: the data-parallel GAN and inference apps after parallel.finish()
mkdir -p gpurun_out/r2x /tmp/r2
O=gpurun_out/r2x
RUN2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632"
timeout 200 $RUN2 -m hypelcnn_b200.gan.gan_train_for_shadow --loader_name SyntheticGULFPORTALTDataLoader \
  --path synthetic:H=64,W=60,samples=600 --gan_type cycle_gan --pairing_method random --batch_size 32 --step 120 \
  --validation_steps 25 --validation_sample_count 50 --base_log_path /tmp/r2/gan > $O/dp_gan.log 2>&1; echo "gan app rc=$?"; grep -E "Output divergence|Best common|Error|error" $O/dp_gan.log | tail -5 | cut -c1-200; ls /tmp/r2/gan*/ | tail -5
