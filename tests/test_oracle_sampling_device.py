"""The warped distribution the device sampling kernel implements (oracle/sampling_device.py::softmax_T: temperature,
top-k, top-p with value-bucket cut-offs) equals what the Hugging Face warpers -- the ones the reference admits,
lade/decoding.py:375-377 -- produce, on rows without tied scores (inside a group of equal scores torch.sort's order, and
with it HF's cut, is undefined)."""
import numpy as np
import pytest
import torch

from oracle import sampling_device as SD


def hf_probs(row, T, k, p):
    from transformers.generation.logits_process import (LogitsProcessorList, TemperatureLogitsWarper, TopKLogitsWarper,
                                                        TopPLogitsWarper)
    ws = LogitsProcessorList()
    if T != 1.0:
        ws.append(TemperatureLogitsWarper(T))
    if k:
        ws.append(TopKLogitsWarper(k))
    if p < 1.0:
        ws.append(TopPLogitsWarper(p))
    scores = ws(None, torch.from_numpy(row.astype(np.float32))[None])
    return torch.softmax(scores.double(), dim=-1)[0].numpy()


@pytest.mark.parametrize("T,k,p", [(1.0, 0, 1.0), (0.7, 0, 1.0), (0.9, 20, 1.0), (1.0, 0, 0.8), (0.8, 50, 0.9),
                                   (1.3, 5, 0.3), (1.0, 1, 1.0), (1.0, 0, 0.05)])
def test_warped_distribution_matches_the_hf_warpers_on_tie_free_rows(T, k, p):
    rng = np.random.default_rng(0)
    for trial in range(4):
        # distinct bf16-representable values in random order (a peaked row: steps of 1/16 up to ~ +-60)
        vals = (np.arange(-1000, 1000, dtype=np.float32) / 16.0)
        row = rng.permutation(vals)[:1500] * (0.2 + 0.3 * trial)
        row = torch.from_numpy(row).to(torch.bfloat16).float().numpy()
        _, counts = np.unique(row, return_counts=True)
        if counts.max() > 1:                       # rounding to bf16 merged two values: make the row tie free again
            row = np.unique(row)
            rng.shuffle(row)
        got = SD.softmax_T(row, T, k, p)
        want = hf_probs(row, T, k, p)
        assert (got > 0).sum() == (want > 0).sum(), (T, k, p, trial)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-12)


def test_ties_are_kept_or_dropped_together():
    row = np.array([3.0, 1.0, 3.0, 2.0, 2.0, 0.0], dtype=np.float32)
    p = SD.softmax_T(row, 1.0, top_k=1)
    assert (p > 0).tolist() == [True, False, True, False, False, False] and abs(p.sum() - 1) < 1e-12
    p = SD.softmax_T(row, 1.0, top_k=3)            # the 3rd largest value is 2.0: both 2.0s stay
    assert (p > 0).tolist() == [True, False, True, True, True, False]
    p = SD.softmax_T(row, 1.0, top_p=0.6)
    assert p[0] == p[2] > 0 and p[5] == 0
