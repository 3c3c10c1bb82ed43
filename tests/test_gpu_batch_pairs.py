"""GPU parity of the batched, device-resident pair entry points (per-pair matching / stereo batches):
orb_stereo_match_batch_device, orb_match_bruteforce_batch_device and orb_search_by_projection_batch_device against the
oracle, pair by pair."""
import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = pytest.mark.gpu
KP = None


def _dev_batch_extract(ex, frames):
    """frames: uint8 [P, h, w] host -> (device tensors kps [P, cap, 7 x f32 as bytes], desc, n) after extract_batch_device"""
    import torch
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    P, h, w = frames.shape
    cap = ex.max_keypoints
    d_img = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros((P, cap, KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((P, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(P, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(d_img.data_ptr(), P, w, h, w, w * h, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    ex.sync()
    return d_img, d_kps, d_desc, d_n, cap


@pytest.mark.parametrize("w,h,nf", [(1241, 376, 2000), (640, 240, 800)])
def test_stereo_match_batch_device(oracle, w, h, nf):
    import torch
    from orb_slam_2_ros_b200 import ORBextractor
    from orb_slam_2_ros_b200._lib import KP_DTYPE
    from orb_slam_2_ros_b200.stereo import ComputeStereoMatchesBatchDevice
    P = 3
    pairs = [synth.synth_stereo_pair(10 + p, w, h)[:2] for p in range(P)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exl, exr = ORBextractor(nf, max_batch=P), ORBextractor(nf, max_batch=P)
    _, kl, dl, nl, cap = _dev_batch_extract(exl, L)
    _, kr, dr, nr, _ = _dev_batch_extract(exr, R)
    bf, fx = np.float32(386.1448), np.float32(718.856)
    b = np.float32(bf / fx)
    ur = torch.zeros((P, cap), dtype=torch.float32, device="cuda"); dep = torch.zeros_like(ur)
    nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    ComputeStereoMatchesBatchDevice(exl, exr, P, kl.data_ptr(), dl.data_ptr(), nl.data_ptr(), kr.data_ptr(), dr.data_ptr(), nr.data_ptr(), cap,
                                    float(bf), float(b), ur.data_ptr(), dep.data_ptr(), nm.data_ptr())
    exl.sync()
    ur, dep, nm, nl_h = ur.cpu().numpy(), dep.cpu().numpy(), nm.cpu().numpy(), nl.cpu().numpy()
    for p in range(P):
        oel, oer = oracle.Extractor(nf), oracle.Extractor(nf)
        okl, odl = oel.extract(L[p]); okr, odr = oer.extract(R[p])
        assert nl_h[p] == len(okl)
        gk = kl[p].cpu().numpy().view(KP_DTYPE).reshape(-1)[:len(okl)]
        assert gk.tobytes() == okl.tobytes()
        kept_o, ur_o, dep_o, _ = oracle.stereo_match(oel, oer, okl, odl, okr, odr, float(bf), float(b))
        assert kept_o > 50
        assert nm[p] == kept_o
        n = len(okl)
        assert np.array_equal(ur[p, :n].view(np.uint32), ur_o.view(np.uint32))
        assert np.array_equal(dep[p, :n].view(np.uint32), dep_o.view(np.uint32))
        assert np.all(ur[p, n:] == -1) and np.all(dep[p, n:] == -1)
