mkdir -p gpurun_out/split
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --light ${CFG:+--config $CFG} 2> gpurun_out/split/$tag.err | tail -1 > gpurun_out/split/$tag.json; python -c "
import json,sys; d=json.load(open('gpurun_out/split/$tag.json')); print('$tag', d['ms_per_step'], d['value'])"; }
run map_split_m2 PN_SM_SPLIT=1 PN_SM_SPLIT_BIAS=-2
run map_split_m4 PN_SM_SPLIT=1 PN_SM_SPLIT_BIAS=-4
run map_split_m6 PN_SM_SPLIT=1 PN_SM_SPLIT_BIAS=-6
run map_split3_w19 PN_SM_SPLIT=1 PN_BWD_STREAMS=3 PN_WGRAD_COST=0.19
run map_split3_w23 PN_SM_SPLIT=1 PN_BWD_STREAMS=3 PN_WGRAD_COST=0.23
run map_split3_w27 PN_SM_SPLIT=1 PN_BWD_STREAMS=3 PN_WGRAD_COST=0.27
CFG=tracking
run trk_split2 PN_SM_SPLIT=1
run trk_split3 PN_SM_SPLIT=1 PN_BWD_STREAMS=3
run trk_base A=1
