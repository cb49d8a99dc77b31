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


def test_cdef_apply_recomputes_directions_when_none_given(b200, refc):
    """svt_b200_cdef_apply_frame_dev with d_dir = d_var = NULL finds the directions itself; the result
    must equal the apply that reuses the arrays of the search."""
    import ctypes as ct
    import torch
    from svt_av1_psy_b200 import dsp
    from svt_av1_psy_b200.pipeline import FramePipeline
    from svt_av1_psy_b200.workload import FrameWorkload
    fp = FramePipeline(FrameWorkload(384, 256), torch)
    fp.step()
    s = torch.cuda.current_stream().cuda_stream
    fp.stage_cdef(s)  # cdef_out without the border extension of the restoration stage
    torch.cuda.synchronize()
    want = fp.cdef_out.clone()
    f = fp.cdef_frame(fp.recon)
    fp.cdef_out.copy_(fp.recon)
    (oy, sy), (ocb, sc), (ocr, _) = fp.plane_views(fp.cdef_out, True)
    rc = dsp.lib.svt_b200_cdef_apply_frame_dev(ct.byref(f), fp.skip.data_ptr(), fp.fb_idx.data_ptr(), fp.app_y.data_ptr(),
                                               fp.app_uv.data_ptr(), None, None, oy, ocb, ocr, sy, sc, s)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(want, fp.cdef_out)
    assert dsp.lib.svt_b200_cdef_apply_frame_dev(ct.byref(f), fp.skip.data_ptr(), fp.fb_idx.data_ptr(), fp.app_y.data_ptr(),
                                                 fp.app_uv.data_ptr(), fp.cdef_dir.data_ptr(), None, oy, ocb, ocr, sy, sc, s) == -4  # SVT_B200_ERR_BAD_ARG
