"""Host-side scheduling logic of the Lasso driver, without a GPU: how a host batch is cut into row chunks and how
iterations are grouped into launches of the resident kernel."""
import types

import pytest
import torch

from decomp_b200 import lasso, ops


@pytest.fixture
def b200(monkeypatch):
    monkeypatch.setattr(torch.cuda, 'get_device_properties',
                        lambda device=None: types.SimpleNamespace(multi_processor_count=148))


def test_row_chunks_are_whole_rounds_of_the_grid(b200):
    assert lasso._row_chunks(1000, 64, 16, None) is None                       # small: one piece
    assert lasso._row_chunks(10000, 1024, 256, None) is None                   # < 4 rounds of 4736 rows
    for B, f, width in [(100000, 1024, 256), (19000, 1024, 256), (400000, 512, 1024), (75777, 256, 32),
                        (3000000, 64, 64)]:
        chunks = lasso._row_chunks(B, f, width, None)
        assert chunks is not None and len(chunks) >= 3
        assert chunks[0][0] == 0 and chunks[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
        unit = 128 * max(1, 148 // -(-width // 64))
        sizes = [r1 - r0 for r0, r1 in chunks]
        assert sizes[0] % unit == 0 and sizes[-1] == unit                      # short first and last chunk
        assert sum(1 for n in sizes if n % unit) <= 1                          # the ragged rest rides in one chunk
        assert sizes[0] <= 2 * unit
        assert max(sizes) * f * 8 <= lasso.PIPELINE_MAX_CHUNK_BYTES + unit * f * 8 or len(chunks) <= 4


def test_pageable_batches_get_uniform_short_chunks(b200):
    chunks = lasso._row_chunks(100000, 1024, 256, None, pinned=False)
    sizes = [r1 - r0 for r0, r1 in chunks]
    assert chunks[0][0] == 0 and chunks[-1][1] == 100000 and all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
    assert sizes[0] == 2 * 4736 and set(sizes[1:-1]) == {4 * 4736} and 4736 <= sizes[-1] < 5 * 4736


def _bare_solver(B, npad, checks, pad=False, n_inplace=10 ** 9, group=None):
    s = object.__new__(lasso.LassoSolver)
    s.resident, s.checks, s.pad, s.B, s.rows_total, s.npad, s.group = True, checks, pad, B, B, npad, group
    s.n_inplace, s.stopped = n_inplace, False
    s.X = torch.zeros((1, 2), dtype=torch.float64)
    s.calls = []
    s._launch_resident = lambda i0, i1: s.calls.append(('resident', i0, i1))
    s._launch = lambda i, out: s.calls.append(('single', i, i + 1))
    return s


def test_launches_without_checks_are_cut_at_the_kernel_limit():
    s = _bare_solver(1000, 32, checks=False)
    s.iterate(0, 70)
    assert s.calls == [('resident', 0, 32), ('resident', 32, 64), ('resident', 64, 70)]
    assert ops.RESIDENT_MAX_ITERS == 32


def test_launches_with_checks_end_on_the_checking_iterations():
    s = _bare_solver(1000, 32, checks=True)
    s.iterate(0, 35)
    # iterations 0, 10, 20, 30 evaluate the test (lasso.py:293/409) and must be the last of their launch
    assert s.calls == [('resident', 0, 1), ('resident', 1, 11), ('resident', 11, 21), ('resident', 21, 31),
                       ('resident', 31, 35)]
    s = _bare_solver(1000, 32, checks=True)
    s.iterate(5, 12)
    assert s.calls == [('resident', 5, 11), ('resident', 11, 12)]


def test_short_launches_of_big_batches_run_per_iteration():
    big = _bare_solver(100000, 256, checks=True)
    big.iterate(0, 21)
    assert big.calls == [('single', 0, 1), ('resident', 1, 11), ('resident', 11, 21)]
    big = _bare_solver(100000, 256, checks=False)
    big.iterate(0, 36)
    assert big.calls == [('resident', 0, 32)] + [('single', i, i + 1) for i in range(32, 36)]
    padded = _bare_solver(100000, 256, checks=True, pad=True)                  # padded buffers: always resident
    padded.iterate(0, 11)
    assert padded.calls == [('resident', 0, 1), ('resident', 1, 11)]


def test_iterate_honours_the_in_place_limit_of_acc_ista():
    s = _bare_solver(1000, 32, checks=False, n_inplace=9)                      # maxiter 10: the last one is finish()'s
    s.iterate(0, 10)
    assert s.calls == [('resident', 0, 9)]


def test_momentum_schedules():
    assert lasso._momentum_schedule('ista', 3) == [0.0, 0.0, 0.0]
    assert lasso._momentum_schedule('acc_ista', 3) == [0.0, 0.25, 0.4]
    m = lasso._momentum_schedule('fista', 3)                                   # (beta - 1) / beta_next, beta_0 = 1
    b1 = 0.5 * (1 + 5 ** 0.5)
    b2 = 0.5 * (1 + (1 + 4 * b1 * b1) ** 0.5)
    assert m[0] == 0.0 and abs(m[1] - (b1 - 1) / b2) < 1e-15


def test_latch_is_polled_after_the_same_iterations_in_both_loops():
    """With a process group, ranks whose shard sizes select different kernel paths must leave the loop together
    (ADVICE r1): both loops read the latch right after checking iterations 50, 100, ... and nowhere else."""
    polled = {}
    for resident in (True, False):
        s = _bare_solver(1000, 32, checks=True)
        s.resident = resident
        reads = []

        class Latch(object):
            def item(self_inner):
                reads.append(s.calls[-1][2] - 1)          # last iteration enqueued when the host reads the latch
                return 0

        s.latch = Latch()
        s.iterate(0, 130)
        polled[resident] = reads
    assert polled[True] == polled[False] == [50, 100]

    # a latch that fires stops both loops after the same iteration
    for resident in (True, False):
        s = _bare_solver(1000, 32, checks=True)
        s.resident = resident

        class Fired(object):
            def item(self_inner):
                return 41

        s.latch = Fired()
        s.iterate(0, 130)
        assert s.stopped and s.calls[-1][2] == 51
        s.iterate(51, 130)                                  # re-entry after a stop enqueues nothing
        assert s.calls[-1][2] == 51


def test_numpy_views_torch_cannot_alias_are_copied():
    import numpy as np
    from decomp_b200 import _device
    a = np.arange(24.0).reshape(4, 6)
    for view in (a[::-1], a[:, ::-1], a[::2, ::3]):
        t = _device._from_numpy(view)
        assert t.shape == view.shape and np.array_equal(t.numpy(), view)
    ro = a.copy()
    ro.flags.writeable = False
    assert np.array_equal(_device._from_numpy(ro).numpy(), a)


def test_sweep_traffic_sums_the_last_sweep(tmp_path):
    """tools/sweep_traffic.py: DRAM bytes of the kernels between the last two normalize_rows launches of an ncu log."""
    import json
    import os
    import subprocess
    import sys
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size",'
            '"Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"']

    def launch(i, name, ns, rd, wr):
        for metric, unit, v in (('dram__bytes_read.sum', 'Mbyte', rd), ('dram__bytes_write.sum', 'byte', wr),
                                ('gpu__time_duration.sum', 'us', ns)):
            rows.append('"%d","1","python","h","%s","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line '
                        'profiler metrics","%s","%s","%s"' % (i, name, metric, unit, v))

    launch(0, 'dcp::normalize_rows_kernel(a)', '4', '0.1', '1,000')
    launch(1, 'void dcp::gemm_f64_kernel<X>(b)', '1,500', '10', '2,000,000')
    launch(2, 'dcp::normalize_rows_kernel(a)', '4', '0', '0')
    launch(3, 'void dcp::gemm_f64_kernel<X>(b)', '2,000', '20.5', '3,000,000')
    launch(4, 'dcp::mu_update_kernel(c)', '10', '1', '500,000')
    launch(5, 'dcp::normalize_rows_kernel(a)', '4', '0', '0')
    log = tmp_path / 'ncu.csv'
    log.write_text('==PROF== Connected\n' + '\n'.join(rows) + '\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'sweep_traffic.py'), str(log)],
                         capture_output=True, text=True, check=True).stdout
    res = json.loads(out)
    assert res['dram_bytes_read'] == pytest.approx(21.5e6) and res['dram_bytes_written'] == pytest.approx(3.5e6)
    assert res['kernel_time_ms'] == pytest.approx(2.014)
    assert [k['name'][:20] for k in res['kernels']] == ['void dcp::gemm_f64_k']      # launches above 0.1 ms only
