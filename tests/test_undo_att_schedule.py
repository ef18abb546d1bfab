"""Subset arithmetic of UNDO_ATTENUATION_AND_OR_PML runs (specfem3d_b200/undo_att.py) against the worked examples in the
reference's own comments (src/specfem3D/compute_add_sources_viscoelastic.F90:771-790, iterate_time_undoatt.F90:88-97)."""
from specfem3d_b200.undo_att import subset_schedule


def _it_tmp_ranges(NSTEP, NT):
    nt, nsub, lengths = subset_schedule(NSTEP, NT)
    out = []
    for iteration_on_subset in range(1, nsub + 1):
        end = NSTEP - (nsub - 1) * nt if iteration_on_subset == 1 else nt  # iterate_time_undoatt.F90:338-346
        its = [NSTEP - (nsub - iteration_on_subset) * nt - k + 1 for k in range(1, end + 1)]
        out.append((its[0], its[-1]))
    return nsub, lengths, out


def test_examples_of_the_reference_comments():
    # NSTEP a multiple of NT_DUMP_ATTENUATION: "1. subset, it_temp goes from 301 down to 1 ... 4. subset, from 1204 down to 904"
    nsub, lengths, r = _it_tmp_ranges(1204, 301)
    assert nsub == 4 and lengths == [301] * 4
    assert r == [(301, 1), (602, 302), (903, 603), (1204, 904)]
    # not a multiple: "1. subset from 298 down to 1, 2. subset from 599 down to 299, 3. subset from 900 down to 600"
    nsub, lengths, r = _it_tmp_ranges(900, 301)
    assert nsub == 3 and lengths == [301, 301, 298]
    assert r == [(298, 1), (599, 299), (900, 600)]


def test_every_forward_step_is_recomputed_once_and_snapshots_fit():
    for NSTEP, NT in ((120, 50), (120, 40), (7, 50), (100, 1), (101, 10)):
        nt, nsub, lengths = subset_schedule(NSTEP, NT)
        assert sum(lengths) == NSTEP and len(lengths) == nsub and all(0 < n <= nt for n in lengths)
        seen = []
        for iteration_on_subset in range(1, nsub + 1):
            end = NSTEP - (nsub - 1) * nt if iteration_on_subset == 1 else nt
            # forward step re-computed at it_of_this_subset = k: NSTEP - it_tmp + 1 (source time (NSTEP - it_tmp) DT - t0)
            steps = [NSTEP - (NSTEP - (nsub - iteration_on_subset) * nt - k + 1) + 1 for k in range(1, end + 1)]
            first_of_frame = (nsub - iteration_on_subset) * nt + 1  # the snapshot was taken before this forward step
            assert steps[0] == first_of_frame and steps == list(range(steps[0], steps[0] + end))
            seen += steps
        assert sorted(seen) == list(range(1, NSTEP + 1))
