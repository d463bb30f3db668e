"""CPU: the caller-side data format of the hot path (`pt/data/common.py:106-180`): the aspect-ratio grouped paired
loader emits exactly the batches the reference's own class emits (tests/golden/pt_reference_loader_golden.json,
made by oracle/make_golden_loader.py from the unmodified reference file), including the items it drops while one
stream's bucket is already full."""
import json
import os

import pytest

from probabilisticteacher_b200.data import AspectRatioGroupedSemiSupDatasetTwoCrop, build_semisup_batch_loader_two_crop

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_loader_golden.json")))


@pytest.mark.parametrize("case", range(len(CASES)))
def test_grouped_paired_loader_matches_the_reference(case):
    c = CASES[case]
    ds = AspectRatioGroupedSemiSupDatasetTwoCrop((c["label"], c["unlabel"]), tuple(c["batch"]))
    got = [[[d["id"] for d in part] for part in b] for b in ds]
    assert got == c["batches"]
    for lq, lk, uq, uk in got:
        assert len(lq) == len(lk) == c["batch"][0] and len(uq) == len(uk) == c["batch"][1]
        assert [i[:-1] for i in lq] == [i[:-1] for i in lk]  # strong / weak views of the same images, same order


def test_batches_are_orientation_pure_and_builder_divides_by_world_size():
    c = CASES[1]
    by_id = {d[0]["id"]: d[0] for d in c["label"] + c["unlabel"]}
    it = build_semisup_batch_loader_two_crop(c["label"], c["unlabel"], 8, 4, world_size=2)  # per rank: (4, 2)
    n = 0
    for lq, lk, uq, uk in it:
        n += 1
        assert len(lq) == 4 and len(uq) == 2
        for grp in (lq, uq):
            assert len({by_id[d["id"]]["width"] > by_id[d["id"]]["height"] for d in grp}) == 1
    assert n == len(c["batches"])
    with pytest.raises(AssertionError):
        build_semisup_batch_loader_two_crop([], [], 3, 4, world_size=2)
