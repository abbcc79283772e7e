set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --batch 4096
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m mdopt_b200.drivers.quantum_surface --lattice_size 3 --bond_dim 64 --error_rate 0.05 --bias_prob 0.1 --num_experiments 101 --error_model Bitflip --seed 123 --num_processes 1 --silent True --tolerance 1e-12 --cut 1e-17 --output_dir gpurun_out
ls gpurun_out/*.pkl
