set -x
python -m pytest tests/test_gpu_parity.py -q -x -k "rasteriser or unfold or sampler_and_composite or full_training or does_not_pin or backward" 2>&1 | tail -5
python -m pytest tests/test_gpu_baseline_shapes.py -q -s -k pretrain 2>&1 | grep "backward vs\|passed\|failed\|Error" | cut -c1-400
mkdir -p /tmp/kp && python - <<'PY'
import numpy as np, json, os
k=np.load('tests/golden/keypoints_body25.npy'); base=json.load(open('tests/golden/keypoints_frame0.json'))
for i,kk in enumerate(k):
    j=json.loads(json.dumps(base)); j["people"][0]["pose_keypoints_2d"]=[float(v) for v in kk.reshape(-1)]
    json.dump(j, open('/tmp/kp/frame%05d_keypoints.json'%i,'w'))
PY
# the reference's launch line (test_start/start.sh) with local paths
( time python3 ./test.py --name d_18Feature_Temporal --checkpoints_dir /tmp/ckpt --pose_path /tmp/kp --pose_tgt_path /tmp/none --use_laplace --bg_path x --texture_path y --TexG part --n_downsample_global 2 --n_blocks_global 10 --ngf_global 48 --use_mask_texture --pose_plus_laplace --n_downsample_bg 2 --n_blocks_bg 2 --no_flip --instance_feat --input_nc 3 --loadSize 512 --resize_or_crop resize --results_dir /tmp/res --which_epoch 30 ) 2>&1 | tail -5
ls /tmp/res | wc -l
python pre_train_tex.py --name t --gpu_ids 0 --batchSize 2 --use_laplace --TexG part --use_mask_texture --n_downsample_global 2 --n_blocks_global 5 --ngf_global 64 --no_flip --instance_feat --input_nc 81 --loadSize 200 --resize_or_crop resize --checkpoints_dir /tmp/ckpt --synthetic_steps 12 2>&1 | tail -3
python train.py --name tr --batchSize 2 --gpu_ids 0 --use_laplace --checkpoints_dir /tmp/ckpt --no_flip --instance_feat --input_nc 3 --loadSize 512 --resize_or_crop resize --lambda_L2 500 --lambda_UV 1000 --lambda_Prob 10 --use_densepose_loss --save_epoch_freq 5 --data_ratio 0.9 --lambda_Temp 500 --synthetic_steps 4 2>&1 | tail -3
python test.py --name tr --checkpoints_dir /tmp/ckpt --pose_path /tmp/kp --use_laplace --pose_plus_laplace --use_mask_texture --input_nc 3 --loadSize 512 --results_dir /tmp/res2 --which_epoch latest --how_many 4 --precision fast 2>&1 | tail -2
