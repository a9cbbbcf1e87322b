"""One unetk_fused_update launch per parameter group (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tiaozhanbei_unet_b200 import ops, train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

dev = torch.device("cuda", 0)
m = AnomalyUNet(3).to(dev).train()
eng = m._engine()
eng.ensure(dev)
opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
upd = opt.fused_update_args(eng)
ops.set_floats(eng.hyper, upd.values)
for rep in range(2):
    for part in (("enc", 3), ("enc", 0)):
        ops.fused_update(eng.update_tables[part], eng.flat_p, eng.flat_g, upd.m, upd.v, eng.gstage, eng.wpack, eng.hyper, upd.mode)
torch.cuda.synchronize()
print("ok")
