#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for cfg in "1" "0"; do
  echo "== tf32 kernels ATTN_PERSIST=$cfg"
  STYLER_ATTN_PERSIST=$cfg timeout 300 python tools/prof_kernels.py --dtype f32 --only ffn1,ffn2_ln,qkv,fc_ln,attention,postnet1,pred_conv_ln,audio_c320_k5,bilstm_h80,groupnorm_relu 2>&1 | tail -12
done
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
from styler_b200 import STYLER, _lib, synthetic as so
dev = torch.device('cuda:0')
m = STYLER(precision='tf32'); m.load_state_dict(so.make_state_dict(0)); m = m.to(dev).eval()
b = so.make_inputs(B=64, L=128, seed=1234, d_mode='const', frames=8)
a = tuple(b[k].to(dev) for k in ("src_seq", "mel_target", "mel_aug", "p_norm", "e_input", "src_len", "mel_len"))
kw = dict(d_target=b["d_target"].to(dev), p_target=b["p_target"].to(dev), e_target=b["e_target"].to(dev), max_src_len=128, max_mel_len=1024, speaker_embed=b["speaker_embed"].to(dev))
def run(tag):
    for _ in range(3): m(*a, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): m(*a, **kw)
    torch.cuda.synchronize(); print(tag, (time.perf_counter() - t0) / 5 * 1e3, 'ms', flush=True)
run('tf32 default')
for name in ('ATTN_PERSIST', 'TC_WIDE', 'TC_2CTA', 'TC_PERSIST'):
    _lib.set_tuning(name, 0); run('tf32 %s=0' % name); _lib.set_tuning(name, -1)
PY
