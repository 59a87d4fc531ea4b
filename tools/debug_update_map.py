"""Debug aid: logs every pair decision of OVO.update_map on the golden replay (compare with tests/golden/update_map.npz 'pairs')."""
import os, sys, tempfile, pathlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gen_golden as GG
from test_gpu_ovo import _build, _replay

U = GG.UPDATE_MAP
tmp = pathlib.Path(tempfile.mkdtemp())
ovo, K, xyz, ids, ins, frames = _build(tmp, extra={"th_centroid": U["th_centroid"], "th_cossim": U["th_cossim"], "th_points": U["th_points"], "log": True})
pins = _replay(ovo, xyz, ids, ins, frames)
before = list(ovo.objects.keys())
ins_in, kfs, drop = GG.update_map_scenario(pins.cpu().numpy(), before)
pairs = []
orig = ovo._same_instance
def logged(a, b, pa, pb):
    r = orig(a, b, pa, pb)
    cen = float(((pa[1] - pb[1]) ** 2).sum().sqrt())
    cos = float(torch.nn.functional.cosine_similarity(a.clip_feature[0], b.clip_feature[0], dim=0))
    pairs.append((a.id, b.id, cen, cos, 1.0 if r else 0.0, tuple(a.clip_feature.shape) == (1, 64), tuple(b.clip_feature.shape) == (1, 64)))
    return r
ovo._same_instance = logged
upd = ovo.update_map((torch.from_numpy(xyz).cuda(), torch.from_numpy(ids).cuda(), torch.from_numpy(ins_in).cuda()), kfs)
np.savez(os.path.join(ROOT, "gpurun_out", "update_map_debug.npz"), pairs=np.array(pairs, np.float64), ins_ids=upd.cpu().numpy(),
         object_ids=np.array(list(ovo.objects.keys())), pins=pins.cpu().numpy())
print("objects", list(ovo.objects.keys()))
