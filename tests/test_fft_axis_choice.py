"""`--fft-axis auto` (snac_b200.param.pick_fft_axis): FFT_AXIS is a build option of the reference; the choice made here must be
a build the reference's own sanity checks accept for the input (chk_grid_fft, src/sanity.f90:66-100; find_fft/eigenvalues,
src/fft.f90:96-189).  Checked on the reference's example decks (read in place: CPU test, this container only) and on the
synthetic decks of the parity tests."""
import os

import pytest

from snac_b200 import decks
from snac_b200.case import setup_case
from snac_b200.param import fft_axes_allowed, pick_fft_axis, read_input

EXAMPLES = "/root/reference/examples"
# as shipped, every example is decomposed 2 x 1 x 2: y is the only direction an FFT build may transform, and it qualifies in
# all of them (periodic or wall-bounded with 'NN' pressure, uniform, the same in every block)
EXPECT = {"_manuscript_taylor_green_vortex": [2], "_manuscript_turbulent_channel": [2], "_manuscript_turbulent_duct": [2],
          "periodic_duct": [2], "_manuscript_lid_driven_cavity": [2], "lid_driven_cavity": [2]}


@pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="reference examples not present")
@pytest.mark.parametrize("name", sorted(EXPECT))
def test_allowed_axes_of_reference_examples(name):
    p = read_input(os.path.join(EXAMPLES, name))
    ok = fft_axes_allowed(p)
    assert ok == EXPECT[name], (name, ok)
    for b in p.blocks:   # undecomposed, the example is free to use any uniform direction with a transform
        b.dims[:] = 1
    wide = fft_axes_allowed(p)
    assert set(ok) <= set(wide) and (name != "_manuscript_taylor_green_vortex" or wide == [1, 2, 3])
    for d in ok:   # never a decomposed, stretched or block-varying direction
        for b in p.blocks:
            assert int(b.dims[d - 1]) == 1 and float(b.gr[d - 1]) == 0.0
    assert pick_fft_axis(p) == max(wide)


def test_synthetic_decks(tmp_path):
    # cavity: stretched in x and z, 'ND' pressure walls in y -> y only (and not decomposed along y when asked so)
    case = decks.make("cavity", str(tmp_path / "a"), fft_axis=0, dims=(2, 1, 1))
    assert fft_axes_allowed(case.params) == [2]
    # the same cavity cut along y: no build accepts it
    case = decks.make("cavity", str(tmp_path / "b"), fft_axis=0, dims=(1, 2, 1))
    assert fft_axes_allowed(case.params) == [] and pick_fft_axis(case.params) == 0
    # L-shaped blocks: only the periodic streamwise direction is common to all blocks
    case = decks.make("lshape", str(tmp_path / "c"), fft_axis=0)
    assert fft_axes_allowed(case.params) == [1]
    # setup_case(..., fft_axis="auto") applies the choice
    case = setup_case(str(tmp_path / "a"), fft_axis="auto")
    assert case.params.fft_axis == 2
