#!/bin/bash
# Runs on the B200 box under gpurun: tests -> logs under gpurun_out/.  Usage: tools/gpu_trip.sh [stage ...]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STAGES="${@:-ops tc e2e bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for s in $STAGES; do
  case $s in
    ops)   timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t_ops.log ;;
    crit)  timeout 900 python -m pytest tests/test_gpu_criterion.py -q -m gpu 2>&1 | tail -60 > gpurun_out/t_crit.log ;;
    opt)  timeout 900 python -m pytest tests/test_gpu_train_step.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/t_opt.log ;;
    optbench*) N=${s#optbench}; N=${N:-1}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_train_step.py > gpurun_out/bench_train_step_$N.log 2>&1 ;;
    bwd)  timeout 1200 python -m pytest tests/test_gpu_backward.py -q -m gpu 2>&1 | tail -120 > gpurun_out/t_bwd.log ;;
    train)  timeout 1500 python -m pytest tests/test_gpu_train.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/t_train.log ;;
    trainbench*) N=${s#trainbench}; N=${N:-1}; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/bench_train.py > gpurun_out/bench_train_$N.log 2>&1 ;;
    ampbench*) N=${s#ampbench}; N=${N:-1}; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_train.py --precision amp > gpurun_out/bench_train_amp_$N.log 2>&1 ;;
    amptests) timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py -q -m gpu -s -k "single_product or amp" 2>&1 | tail -40 > gpurun_out/t_amp.log ;;
    benchnopdl) FB200_TC_PDL=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-other-configs --no-train-leg --no-cpu-baseline > gpurun_out/bench_nopdl.log 2> gpurun_out/bench_nopdl.err ;;
    benchpdl) timeout 900 python bench.py --steps 10 --warmup 3 --no-other-configs --no-train-leg --no-cpu-baseline > gpurun_out/bench_pdl.log 2> gpurun_out/bench_pdl.err ;;
    mfbench32) FB200_BENCH_PRECISION=fp32_tc FB200_BENCH_PARITY_MODE=0 timeout 900 python tools/bench_mf.py > gpurun_out/bench_mf_fp32_tc.txt 2>&1 ;;
    bisebench32) FB200_BENCH_PRECISION=fp32_tc FB200_BENCH_PARITY_MODE=0 timeout 900 python tools/bench_bisenet.py > gpurun_out/bench_bisenet_fp32_tc.txt 2>&1 ;;
    sanitize) (timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mf.py -q -m gpu -x -k "masked_attention_split or mask_build" 2>&1 | grep -v "Host Frame" | tail -40; echo "rc=$?"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_backward.py -q -m gpu -x -k "batchnorm or layernorm or attention_grads or pool" 2>&1 | grep -v "Host Frame" | tail -40; echo "rc=$?") > gpurun_out/sanitize.log 2>&1 ;;
    ncu_stem) timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_conv_tiled -c 1 -o gpurun_out/prof_stem python tools/profile_step.py fp32_tc > gpurun_out/ncu_stem.log 2>&1 ;;
    stem2) FB200_STEM_TILED=2 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_stem2.csv python tools/profile_step.py fp32_tc > gpurun_out/ncu_list_stem2.log 2>&1; FB200_STEM_TILED=2 timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k stem 2>&1 | tail -3 > gpurun_out/t_stem2.log ;;
    headtrace) timeout 300 python tools/head_trace.py > gpurun_out/head_trace.txt 2>&1 ;;
    mattnqb) (for qb in 32 48 64 112; do echo "### FB200_MATTN_QB=$qb"; FB200_MATTN_QB=$qb FB200_BENCH_PRECISION=fp32_tc FB200_BENCH_PARITY_MODE=0 timeout 300 python tools/bench_mf.py | grep -E 'images_per_s|attention_masked'; done; true) > gpurun_out/mattn_qb.txt 2>&1 ;;
    wgmicro) python tools/wgrad_micro.py > gpurun_out/wgrad_micro.txt 2>&1 ;;
    wgncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 3 -c 1 -o gpurun_out/prof_wgrad python tools/wgrad_micro.py fpn_rep_3x3_80 > gpurun_out/wgrad_ncu.log 2>&1 ;;
    microres) (python tools/conv_micro.py s0_2c_res s1_2c_res s2_2c_res; FB200_TC_RES_TMA=0 python tools/conv_micro.py s0_2c_res s1_2c_res s2_2c_res; true) > gpurun_out/conv_micro_res.txt 2>&1 ;;
    tc)    timeout 900 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu 2>&1 | tail -80 > gpurun_out/t_tc.log ;;
    e2e)   timeout 1200 python -m pytest tests/test_gpu_e2e.py -q -m gpu 2>&1 | tail -60 > gpurun_out/t_e2e.log ;;
    smoke) timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 --no-other-configs > gpurun_out/bench.log 2> gpurun_out/bench.err ;;
    benchfull) (time timeout 900 python bench.py) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err ;;
    bench32) timeout 900 python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.log 2> gpurun_out/bench_fp32.err; timeout 900 python bench.py --steps 10 --warmup 3 --precision fp32_tc --no-cpu-baseline > gpurun_out/bench_fp32tc.log 2> gpurun_out/bench_fp32tc.err ;;
    budget) timeout 1200 python tools/error_budget.py > gpurun_out/error_budget.txt 2>&1 ;;
    micro_cfg) for c in 0 1; do (FB200_TC_CFG=$c timeout 300 python tools/conv_micro.py rep_3x3_80 rep_3x3_40 s2_2b s3_2b csp_1x1_80; true) > gpurun_out/conv_micro_cfg$c.txt 2>&1; done ;;
    ncu_pair) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_pair python tools/conv_micro.py rep_3x3_80 > gpurun_out/ncu_pair.log 2>&1 ;;
    trace2) (FB200_TC_CTA2=0 timeout 300 python tools/conv_trace.py rep_3x3_80 s2_2a; FB200_TC_CTA2=0 FB200_GRID_CAP=74 timeout 300 python tools/conv_trace.py rep_3x3_80; FB200_TC_CTA2=0 FB200_GRID_CAP=16 timeout 300 python tools/conv_trace.py rep_3x3_80; timeout 300 python tools/conv_trace.py rep_3x3_80; true) 2>&1 | grep -v "tile[2-9]:\|tile1[0-9]:" > gpurun_out/conv_trace2.txt ;;
    trace3) (FB200_TC_DBG=8 FB200_TC_CTA2=0 timeout 300 python tools/conv_trace.py rep_3x3_80; FB200_TC_DBG=8 timeout 300 python tools/conv_trace.py rep_3x3_80; FB200_TC_DBG=8 FB200_TC_CTA2=0 FB200_TC_BN=128 timeout 300 python tools/conv_trace.py rep_3x3_80;  FB200_TC_DBG=8 FB200_TC_CTA2=0 FB200_TC_BN=64 timeout 300 python tools/conv_trace.py rep_3x3_80; true) 2>&1 | grep -v "tile[2-9]:\|tile1[0-9]:" > gpurun_out/conv_trace3.txt ;;
    trace4) (for d in 8 12 28 24; do for c in 0 1; do echo "### DBG=$d CTA2=$c"; FB200_TC_DBG=$d FB200_TC_CTA2=$c timeout 300 python tools/conv_trace.py rep_3x3_80; done; done; echo "### DBG=12 BN=128"; FB200_TC_DBG=12 FB200_TC_CTA2=0 FB200_TC_BN=128 timeout 300 python tools/conv_trace.py rep_3x3_80; echo "### DBG=28 BN=128"; FB200_TC_DBG=28 FB200_TC_CTA2=0 FB200_TC_BN=128 timeout 300 python tools/conv_trace.py rep_3x3_80; echo "### DBG=4 (TMA on, epilogue off)"; FB200_TC_DBG=4 FB200_TC_CTA2=0 timeout 300 python tools/conv_trace.py rep_3x3_80; FB200_TC_DBG=4 FB200_TC_CTA2=1 timeout 300 python tools/conv_trace.py rep_3x3_80; true) 2>&1 | grep -v "^cta\|tile[0-9]" > gpurun_out/conv_trace4.txt ;;
    trace5) (for d in 36 4; do for c in 0 1; do echo "### DBG=$d CTA2=$c"; FB200_TC_DBG=$d FB200_TC_CTA2=$c timeout 300 python tools/conv_trace.py rep_3x3_80 s2_2a; done; done; true) 2>&1 | grep -v "^cta\|tile[0-9]" > gpurun_out/conv_trace5.txt ;;
    trace6) (for d in 0 64 192; do echo "### DBG=$d CTA2=0"; FB200_TC_DBG=$d FB200_TC_CTA2=0 timeout 300 python tools/conv_trace.py rep_3x3_80; done; true) 2>&1 | grep -v "^cta\|tile[0-9]" > gpurun_out/conv_trace6.txt ;;
    microfs) (FB200_TC_FS=0 timeout 300 python tools/conv_micro.py --split; true) > gpurun_out/conv_micro_split_seg.txt 2>&1; (timeout 300 python tools/conv_micro.py --split; true) > gpurun_out/conv_micro_split_fs.txt 2>&1 ;;
    benchfp16) timeout 900 python bench.py --steps 20 --warmup 3 --precision fp16 --quick --no-cpu-baseline > gpurun_out/bench_fp16.log 2> gpurun_out/bench_fp16.err ;;
    ncu_fs) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_fs python tools/conv_micro.py --split rep_3x3_80 > gpurun_out/ncu_fs.log 2>&1 ;;
    alltests) timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t_all.log ;;
    ablib) timeout 900 python tools/ab_lib.py focoos_b200/lib/libfocoos_b200_r01.so $(ls focoos_b200/lib/libfocoos_b200_*.so | grep -v r01) focoos_b200/lib/libfocoos_b200.so > gpurun_out/ab_lib.txt 2>&1 ;;
    ddp2) timeout 1500 python -m pytest tests/test_gpu_train_ddp.py -q -m gpu -s 2>&1 | tail -15 > gpurun_out/t_ddp2.log ;;
    bench2) timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-other-configs > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err ;;
    bench4) timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 20 --warmup 3 --no-other-configs > gpurun_out/bench_4gpu.log 2> gpurun_out/bench_4gpu.err ;;
    bench8) timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 3 --no-other-configs > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err ;;
    benchdefault) (time timeout 1500 python bench.py) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err ;;
    ncu_list32) timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_fp32_tc.csv python tools/profile_step.py fp32_tc > gpurun_out/ncu_list_fp32_tc.log 2>&1 ;;
    trace) (FB200_TC_CTA2=0 timeout 300 python tools/conv_trace.py rep_3x3_80 rep_3x3_40 s3_2b s2_2a s0_2c; timeout 300 python tools/conv_trace.py rep_3x3_80 rep_3x3_40; true) > gpurun_out/conv_trace.txt 2>&1 ;;
    budget2) timeout 1200 python tools/error_budget.py tc:3323 tc:3331 tc:3332 tc:2222 tc:1111 > gpurun_out/error_budget2.txt 2>&1 ;;
    micro01) (FB200_TC_CTA2=0 timeout 300 python tools/conv_micro.py; true) > gpurun_out/conv_micro_cta1.txt 2>&1; (timeout 300 python tools/conv_micro.py; true) > gpurun_out/conv_micro_cta2.txt 2>&1 ;;
    layers) timeout 600 python tools/layer_roofline.py 32 fp32_tc > gpurun_out/layer_roofline_fp32_tc.txt 2>&1; timeout 600 python tools/layer_roofline.py 32 fp16 > gpurun_out/layer_roofline_fp16.txt 2>&1 ;;
    micropair) (for nc in 0 1; do echo "### FB200_TC_NCAT=$nc"; FB200_TC_NCAT=$nc timeout 300 python tools/conv_micro.py --pair stem2 stem3 s0_2b s1_2b s0_2a s0_2c s1_2a s2_2b; done; true) > gpurun_out/conv_micro_pair.txt 2>&1 ;;
    attnqb) (for qb in 64 96 128 160 192; do echo "### FB200_ATTN_QB=$qb"; FB200_ATTN_QB=$qb timeout 300 python tools/head_micro.py attn; done; true) > gpurun_out/attn_qb.txt 2>&1 ;;
    headmicro) (timeout 600 python tools/head_micro.py; true) > gpurun_out/head_micro.txt 2>&1 ;;
    micro) (python tools/conv_micro.py; true) > gpurun_out/conv_micro.txt 2>&1 ;;
    micro_ncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 2 -o gpurun_out/prof_micro python tools/conv_micro.py rep_3x3_80 s0_2c_res > gpurun_out/micro_ncu.log 2>&1 ;;
    mf)    timeout 900 python -m pytest tests/test_gpu_mf.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_mf.log ;;
    mfbench) timeout 900 python tools/bench_mf.py > gpurun_out/bench_mf.txt 2>&1 ;;
    bise)  timeout 900 python -m pytest tests/test_gpu_bisenet.py -q -m gpu 2>&1 | tail -40 > gpurun_out/t_bise.log ;;
    bisebench) timeout 900 python tools/bench_bisenet.py > gpurun_out/bench_bisenet.txt 2>&1 ;;
    ref)   timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1 ;;
    ncu_list) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 940 -c 240 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1 ;;
    list_*) P=${s#list_}; timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$P.csv python tools/profile_step.py $P > gpurun_out/ncu_list_$P.log 2>&1 ;;
    ncu_full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 8 -o gpurun_out/prof_conv_tc python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ;;
  esac
  echo "stage $s exit $?" >> gpurun_out/stages.txt
done
tail -5 gpurun_out/t_*.log gpurun_out/bench.log 2>/dev/null | tail -60
