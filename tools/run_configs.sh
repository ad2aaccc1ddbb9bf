#!/bin/bash
# Driver-style lines for every BASELINE config at N GPUs of one box (run under `gpurun --gpus N`):
#   config 2 + 3: bench.py (N = 1: certification batches; N > 1: whole certifications, samples sharded + NCCL int64 sum)
#   config 4:     bench.py --frm facenet
#   config 5:     (i) gallery builder (embedding extraction, identities sharded), (ii) 1 M-row gallery sharded over the ranks
# usage: tools/run_configs.sh N [tag]   -> gpurun_out/<tag>_N<N>_*.json
N=${1:-1}; TAG=${2:-cfg_r02}; OUT=gpurun_out; mkdir -p $OUT
if [ "$N" -gt 1 ]; then RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; else RUN="python"; fi
timeout 400 $RUN bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_N${N}_bench.json 2> $OUT/${TAG}_N${N}_bench.err
timeout 400 $RUN bench.py --gpus $N --steps 8 --warmup 3 --frm facenet > $OUT/${TAG}_N${N}_facenet.json 2> $OUT/${TAG}_N${N}_facenet.err
timeout 300 $RUN tools/build_gallery.py --identities $((2048 * N)) --chunk 128 > $OUT/${TAG}_N${N}_gallery_build.json 2> $OUT/${TAG}_N${N}_gallery_build.err
timeout 300 $RUN tools/bench_gallery_sharded.py --rows 1000000 > $OUT/${TAG}_N${N}_gallery_match.json 2> $OUT/${TAG}_N${N}_gallery_match.err
for f in bench facenet gallery_build gallery_match; do echo "== $f"; tail -1 $OUT/${TAG}_N${N}_$f.json | cut -c1-260; done
