#!/bin/bash
# Round 2, eight GPUs (gpurun --gpus 8): headline bench, C3, whole-scene inference and the GAN bench at N = 8; the
# data-parallel GAN and classifier apps at N = 2 (after the device-selection fix in gan_train_for_shadow)
mkdir -p gpurun_out/r2y /tmp/r2
O=gpurun_out/r2y
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621"
RUN2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622"
timeout 600 $RUN8 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu.log 2>&1; tail -1 $O/bench_8gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=8', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('final_loss'))"
timeout 600 $RUN8 bench.py --gpus 8 --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51_8gpu.log 2>&1; tail -1 $O/bench_c3_51_8gpu.log | cut -c1-300
timeout 600 $RUN8 scripts/bench_inference.py > $O/inference_8gpu.json 2> $O/inference_8gpu.err; tail -1 $O/inference_8gpu.json | cut -c1-300; tail -2 $O/inference_8gpu.err
timeout 300 $RUN8 scripts/bench_gan.py --batches 32,16384 > $O/gan_8gpu.json 2> $O/gan_8gpu.err; cut -c1-200 $O/gan_8gpu.json; tail -2 $O/gan_8gpu.err
cat > /tmp/r2/alg.json <<'JSON'
{"batch_size": 48, "drop_out_ratio": 0.70, "filter_count": 64, "learning_rate": 0.0003, "learning_rate_decay_factor": 0.96,
 "learning_rate_decay_step": 350, "lrelu_alpha": 0.18, "optimizer": "AdamOptimizer", "bn_decay": 0.95,
 "l2regularizer_scale": 0.00001, "spectral_hierarchy_level": 3, "spatial_hierarchy_level": 3, "degradation_coeff": 3,
 "use_residual": true}
JSON
timeout 300 $RUN2 -m hypelcnn_b200.gan.gan_train_for_shadow --loader_name SyntheticGULFPORTALTDataLoader \
  --path synthetic:H=64,W=60,samples=600 --gan_type cycle_gan --pairing_method random --batch_size 32 --step 120 \
  --validation_steps 25 --validation_sample_count 50 --base_log_path /tmp/r2/gan > $O/dp_gan.log 2>&1; tail -3 $O/dp_gan.log | cut -c1-200
timeout 300 $RUN2 -m hypelcnn_b200.classify.train_for_classification --loader_name SyntheticGRSS2013DataLoader \
  --path synthetic:H=40,W=60,samples=800 --neighborhood 3 --train_ratio 1.0 --test_ratio 0.1 --batch_size 64 --step 40 \
  --algorithm_param_path /tmp/r2/alg.json --perform_validation True --validation_steps 15 --save_checkpoint_steps 20 \
  --base_log_path /tmp/r2/classify > $O/dp_classify.log 2>&1; tail -3 $O/dp_classify.log | cut -c1-200
timeout 600 $RUN2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_2gpu.log 2>&1; tail -1 $O/bench_2gpu.log | cut -c1-250
