"""GPU parity at picture scale: one frame through the T2 pipeline (the path bench.py times) against
the reference arm -- the reference's own kernels driven by oracle/ref_driver.c -- bit for bit, at a
small size and at BASELINE's 1920x1080."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(384, 256), (1920, 1080)])
def test_frame_pipeline_matches_reference(b200, refc, size):
    import torch
    import bench
    from svt_av1_psy_b200.pipeline import FramePipeline
    from svt_av1_psy_b200.workload import FrameWorkload
    fp = FramePipeline(FrameWorkload(*size), torch)
    bench.check_against_reference(fp, torch)
    # size-independent property: a second pass over the same inputs is idempotent
    a = fp.final.clone(), fp.qcoeff.clone(), fp.me_mv.clone()
    fp.step()
    torch.cuda.synchronize()
    assert torch.equal(a[0], fp.final) and torch.equal(a[1], fp.qcoeff) and torch.equal(a[2], fp.me_mv)
