"""Input staging (fact_clip_b200/staging.py): same values as the reference's load_feature (np.load -> optional .T ->
astype(float32), utils/dataset.py:12-21), batch composition, arena ring recycling.  Host logic only: runs without a GPU."""
import os

import numpy as np
import pytest
import torch

from fact_clip_b200 import staging as S


def reference_load(feature_dir, video, transpose):
    """What utils/dataset.py:12-21 returns (restated; the reference itself needs its package layout)."""
    feature = np.load(os.path.join(feature_dir, video + '.npy'))
    if transpose:
        feature = feature.T
    if feature.dtype != np.float32:
        feature = feature.astype(np.float32)
    return feature


@pytest.fixture
def feature_dir(tmp_path):
    rng = np.random.default_rng(0)
    lens = dict(a=37, b=5, c=64, d=1, e=20)
    for name, T in lens.items():
        np.save(tmp_path / f'{name}.npy', rng.standard_normal((24, T)))                 # (D, T) float64: transpose + cast
        np.save(tmp_path / f'{name}_td.npy', rng.standard_normal((T, 24)).astype(np.float32))
    return str(tmp_path), lens


@pytest.mark.parametrize('transpose', [True, False])
def test_values_match_reference_loader(feature_dir, transpose):
    d, lens = feature_dir
    names = list(lens) if transpose else [n + '_td' for n in lens]
    st = S.FeatureStager(d, names, transpose=transpose, batch_videos=2, pin=False, workers=3)
    seen = []
    for batch in st:
        for n, x in zip(batch.names, batch.seqs):
            ref = reference_load(d, n, transpose)
            assert x.dtype == torch.float32 and tuple(x.shape) == ref.shape and x.is_contiguous()
            assert np.array_equal(x.numpy(), ref)
            assert np.array_equal(S.load_feature(d, n, transpose), ref)
        seen += batch.names
        batch.release()
    assert seen == names and len(st) == 3 and st.frames() == sum(lens.values()) and st.dim == 24
    st.close()


def test_bf16_arena_and_length_sorted_batches(feature_dir):
    d, lens = feature_dir
    st = S.FeatureStager(d, list(lens), transpose=True, batch_videos=2, dtype=torch.bfloat16, pin=False, sort_by_length=True)
    order = []
    for batch in st:
        for n, x in zip(batch.names, batch.seqs):
            assert x.dtype == torch.bfloat16
            assert torch.equal(x, torch.from_numpy(reference_load(d, n, True)).to(torch.bfloat16))
        order += batch.names
        batch.release()
    assert [lens[n] for n in order] == sorted(lens.values())


def test_arena_ring_blocks_until_release(feature_dir):
    """depth arenas: the producer cannot overwrite a batch the consumer still holds."""
    d, lens = feature_dir
    st = S.FeatureStager(d, list(lens), transpose=True, batch_videos=1, pin=False, depth=2)
    it = iter(st)
    b0, b1 = next(it), next(it)
    keep0 = b0.seqs[0].clone()
    # both arenas are held: nothing may be staged into them until one is released
    import time
    time.sleep(0.2)
    assert torch.equal(b0.seqs[0], keep0)
    b0.release()
    b2 = next(it)                                   # re-uses b0's arena
    assert b2.seqs[0].data_ptr() == b0.seqs[0].data_ptr()
    assert np.array_equal(b1.seqs[0].numpy(), reference_load(d, b1.names[0], True))
    b1.release(); b2.release()
    rest = [b.names[0] for b in it]
    assert rest == list(lens)[3:]


def test_loader_errors_surface(feature_dir):
    d, lens = feature_dir
    np.save(os.path.join(d, 'bad.npy'), np.zeros((3, 4, 5)))
    with pytest.raises(AssertionError, match='2-D'):
        S.FeatureStager(d, ['a', 'bad'], transpose=True, pin=False)
    np.save(os.path.join(d, 'wide.npy'), np.zeros((30, 7)))
    with pytest.raises(AssertionError, match='feature dimension'):
        S.FeatureStager(d, ['a', 'wide'], transpose=True, pin=False)


def test_device_transpose_mode_keeps_file_layout(feature_dir):
    d, lens = feature_dir
    st = S.FeatureStager(d, list(lens), transpose=True, batch_videos=3, pin=False, device_transpose=True)
    for batch in st:
        assert batch.channel_major
        for n, x in zip(batch.names, batch.seqs):
            assert tuple(x.shape) == (24, lens[n]) and x.is_contiguous() and x.dtype == torch.float32
            assert np.array_equal(x.numpy().T, reference_load(d, n, True))          # the float64 file is cast on the way in
        batch.release()
    with pytest.raises(AssertionError):
        S.FeatureStager(d, list(lens), transpose=False, pin=False, device_transpose=True)
