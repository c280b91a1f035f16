#!/bin/bash
# CIFAR-10 full: 60 000 iterations at lr 1e-3, then 5 000 at 1e-4 and 5 000 at 1e-5, each stage resuming from the
# previous stage's snapshot (reference: examples/cifar10/cifar10_full_solver{,_lr1,_lr2}.prototxt).
#   examples/cifar10/train_full.sh [NUM_GPUS] [extra caffe_main flags]
set -e
cd "$(dirname "$0")/../.."
N=${1:-1}
python -m poseidon_b200.models.zoo --out models --only cifar10_full
[ -d data/cifar10 ] && [ ! -d examples/cifar10/cifar10_train_leveldb ] && bash examples/cifar10/create_cifar10.sh
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29400 \
    -m poseidon_b200.tools.caffe_main train --net_outputs=output/cifar10_full "$@"
}
run --solver=models/cifar10_full/solver.prototxt "${@:2}"
run --solver=models/cifar10_full/solver_lr1.prototxt --snapshot=cifar10_full_iter_60000.solverstate "${@:2}"
run --solver=models/cifar10_full/solver_lr2.prototxt --snapshot=cifar10_full_iter_65000.solverstate "${@:2}"
