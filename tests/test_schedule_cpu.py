"""Host logic of the decoder's wavefront schedules (rsis_b200/modules/model.py::wavefront_schedule, run_wavefront): every
cell of the (level, step) loop nest of /root/reference/src/test.py:37-44 x src/modules/model.py:129-165 runs exactly once,
after the cells it depends on, and the ping-pong buffers the schedule relies on are never overwritten while a reader is
still due.  Pure Python: no GPU, no library."""
import pytest

from rsis_b200.modules.model import wavefront_schedule


@pytest.mark.parametrize("T,nlev", [(1, 5), (2, 5), (5, 5), (10, 5), (16, 5), (20, 5), (3, 1), (7, 2)])
@pytest.mark.parametrize("skew", [1, 2])
def test_every_cell_once_and_after_its_dependencies(T, nlev, skew):
    waves = wavefront_schedule(T, nlev, skew)
    when = {}
    for w, wave in enumerate(waves):
        assert wave or (T == 1 and skew == 2), "no empty launches (T = 1 skewed: every other one, which run_wavefront skips)"
        for cell in wave:
            assert cell not in when
            when[cell] = w
    assert set(when) == {(l, t) for l in range(nlev) for t in range(T)}
    assert len(waves) == T + skew * (nlev - 1)
    for (l, t), w in when.items():
        if t > 0:
            assert when[(l, t - 1)] < w                      # its own state (clstm.py:43: prev_hidden, prev_cell)
        if l > 0:
            # the x2 upsampling of (l-1, t) runs AFTER that cell's launch and BEFORE this one; with skew 2 it has the
            # launch in between to itself (side stream), with skew 1 it sits between the two launches
            assert w - when[(l - 1, t)] == skew


@pytest.mark.parametrize("T", [1, 2, 3, 10, 16])
def test_skewed_schedule_buffer_hazards(T):
    """skew 2, as run_wavefront uses it.  Buffers (DecoderWorkspace): X[l][p] = [up(h_{l-1}) | h_prev_l] ping-pong by
    step parity; h2[l][p] float32 hidden state of level l < nlev-1, ping-pong by step parity.  Timeline unit: launch w
    occupies [w, w + 1); the upsampling of the cells of launch w occupies (w + 1, w + 2) -- it starts when launch w has
    finished and launch w + 2 waits for it."""
    nlev = 5
    when = {c: w for w, wave in enumerate(wavefront_schedule(T, nlev, 2)) for c in wave}
    for (l, t), w in when.items():
        p = t & 1
        # (1) cell (l, t) writes h16 into the h_prev part of X[l][1 - p]; its last reader was cell (l, t - 1) [buffer
        #     parity (t - 1) & 1 == 1 - p], its next reader is cell (l, t + 1)
        if t > 0:
            assert when[(l, t - 1)] < w
        if t + 1 < T:
            assert when[(l, t + 1)] > w
        if l + 1 < nlev:
            up_begin, up_end = w + 1, w + 2          # the upsampling of (l, t): reads h2[l][p], writes up part of X[l+1][p]
            # (2) h2[l][p] is overwritten next by cell (l, t + 2): not before the upsampling has read it
            if t + 2 < T:
                assert when[(l, t + 2)] >= up_end
            # (3) the up part of X[l+1][p] was last read by cell (l + 1, t - 2): finished before the upsampling starts
            if t >= 2:
                assert when[(l + 1, t - 2)] + 1 <= up_begin
            # (4) ... and its consumer, cell (l + 1, t), starts when the upsampling has ended
            assert when[(l + 1, t)] >= up_end
