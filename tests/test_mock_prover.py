"""MockProver mirror (spectre_b200/mock.py): the synthetic circuits are satisfied; a broken gate, lookup or copy is named."""
import pytest

from spectre_b200 import circuits, mock, plonk


@pytest.mark.parametrize("shape", ["aggregation", "wide", "halo2lib"])
def test_synthetic_witnesses_are_satisfied(shape):
    k, inst = 7, [5, 6, 7]
    if shape == "aggregation":
        cs = circuits.aggregation_shape(); fixed, adv, copies = circuits.aggregation_witness(cs, k, inst, 3, 20); adv = [adv]
    elif shape == "wide":
        cs = circuits.wide_shape(3); fixed, adv, copies = circuits.wide_witness(cs, k, inst, 3, 20)
    else:
        cs = circuits.halo2lib_shape(3, 2); fixed, adv, copies = circuits.halo2lib_witness(cs, k, inst, 3, 20, num_gate_advice=3, num_lookup_advice=2)
    mock.assert_satisfied(cs, k, fixed, adv, [inst], copies)


def test_failures_are_located():
    k, inst = 6, [9]
    cs = circuits.aggregation_shape()
    fixed, adv, copies = circuits.aggregation_witness(cs, k, inst, 3, 8)
    bad = adv.copy(); bad[7] = plonk.fr_mont(1)                 # d of the second group
    assert any("gate 0 not satisfied on row 4" in f for f in mock.run(cs, k, fixed, [bad], [inst], copies))
    bad = adv.copy(); bad[8] = plonk.fr_mont(99)                # looked-up a of the third group (table is [0, 8))
    assert any("lookup 0: input on row 8" in f for f in mock.run(cs, k, fixed, [bad], [inst], copies))
    with pytest.raises(mock.VerifyFailure, match="copy constraint"):
        mock.assert_satisfied(cs, k, fixed, [adv], [[10]], copies)   # the public input no longer equals the copied cell
