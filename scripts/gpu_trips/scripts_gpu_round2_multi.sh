#!/bin/bash
# Round 2, two GPUs (gpurun --gpus 2 --timeout 600 -- 'bash scripts_gpu_round2_multi.sh'): the app-level data-parallel
# paths wired late in round 1 (never run on more than one GPU so far), then the 2-GPU headline.
mkdir -p gpurun_out /tmp/r2
cat > /tmp/r2/alg.json <<'JSON'
{"batch_size": 48, "drop_out_ratio": 0.70, "filter_count": 64, "learning_rate": 0.0003, "learning_rate_decay_factor": 0.96,
 "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "optimizer": "AdamOptimizer", "bn_decay": 0.95,
 "l2regularizer_scale": 0.00001, "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3, "degradation_coeff": 3,
 "use_residual": true}
JSON
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
# classifier: strided training split, overlapped all-reduce, rank 0 writes checkpoints / summaries
timeout 300 $RUN -m hypelcnn_b200.classify.train_for_classification --loader_name SyntheticGRSS2013DataLoader \
  --path synthetic:H=40,W=60,samples=800 --neighborhood 3 --train_ratio 1.0 --test_ratio 0.1 --batch_size 64 --step 40 \
  --algorithm_param_path /tmp/r2/alg.json --perform_validation True --validation_steps 15 --save_checkpoint_steps 20 \
  --base_log_path /tmp/r2/classify > gpurun_out/dp_classify.log 2>&1; tail -4 gpurun_out/dp_classify.log
# whole-scene inference: contiguous pixel slices per rank, one MIN all-reduce of the class image
timeout 300 $RUN -m hypelcnn_b200.classify.infer_for_classification --loader_name SyntheticGRSS2013DataLoader \
  --path synthetic:H=40,W=60,samples=800 --neighborhood 3 --batch_size 256 --algorithm_param_path /tmp/r2/alg.json \
  --base_log_path /tmp/r2/classify/syntheticgrss2013ldr_hypelcnnmdl_trn100_alg_7x7 --output_path /tmp/r2 \
  > gpurun_out/dp_infer.log 2>&1; tail -2 gpurun_out/dp_infer.log
# GAN: pair rows strided with equal counts, one all-reduce per train op, chief-only validation
timeout 300 $RUN -m hypelcnn_b200.gan.gan_train_for_shadow --loader_name SyntheticGULFPORTALTDataLoader \
  --path synthetic:H=64,W=60,samples=600 --gan_type dcl_gan --pairing_method random --batch_size 32 --step 120 \
  --validation_steps 25 --validation_sample_count 50 --base_log_path /tmp/r2/gan > gpurun_out/dp_gan.log 2>&1; tail -3 gpurun_out/dp_gan.log
# headline at N = 2
timeout 600 $RUN bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log
