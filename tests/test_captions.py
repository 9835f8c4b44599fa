"""Caption assembly and result files and the training-loop surface that needs no GPU (f4):
Vocabulary.get_sentence, the eval json / test csv writers, the staircase learning-rate schedule."""
import csv
import json
import os

import numpy as np

import sat_b200
from oracle import ref_step as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def vocab():
    # the first 120 entries of the reference's data/vocabulary.csv (same file format: Vocabulary.save, vocabulary.py:65-70)
    return sat_b200.Vocabulary(save_file=os.path.join(GOLD, "vocabulary_head.csv"))


def test_vocabulary_file_and_special_ids():
    v = vocab()
    assert v.words[0] == "<start>" and v.words[1] == "a" and v.words[2] == "."
    assert v.eos_id == 2 and v.word2idx["a"] == 1 and v.size == 120


def test_get_sentence_equals_the_reference_restatement():
    v = vocab()
    rng = np.random.RandomState(0)
    apos = [i for i, w in enumerate(v.words) if w.startswith("'")]
    punct = [i for i, w in enumerate(v.words) if w in (",", ".", "'s")]
    for trial in range(200):
        n = rng.randint(1, 12)
        ids = [int(x) for x in rng.randint(1, v.size, n)]
        if trial % 3 == 0:
            ids.insert(rng.randint(0, n + 1), 2)              # a '.' somewhere: the caption stops there
        if trial % 5 == 0 and punct:
            ids.insert(rng.randint(0, len(ids) + 1), punct[rng.randint(len(punct))])
        if trial % 7 == 0 and apos:
            ids.insert(rng.randint(1, len(ids) + 1), apos[0])
        got = v.get_sentence(ids)
        assert got == R.get_sentence(v.words, ids)
        assert got.endswith(".") and got.count(".") == 1
    assert v.get_sentence([1, v.word2idx["man"], v.word2idx["with"], 1, v.word2idx["dog"]]) == "a man with a dog."


def test_result_writers_match_the_reference_formats(tmp_path):
    import pandas as pd
    v = vocab()
    C = sat_b200.model.CaptionData
    beams = [[C([1, 10, 2], 0.25, True), C([1, 11, 2], 0.2, True)], [C([5, 12, 13], 1e-3, False)], [C([1, 2], 0.5, True)]]
    captions, scores = sat_b200.assemble_captions(beams, v, fake_count=1)            # the last entry pads the batch
    assert len(captions) == 2 and captions[0] == v.get_sentence([1, 10, 2]) and scores == [0.25, 1e-3]
    # eval: json list of {"image_id", "caption"} (base_model.py:87-88, 109-111)
    p = str(tmp_path / "results.json")
    sat_b200.write_eval_results(p, np.array([391895, 522418]), captions)
    assert json.load(open(p)) == [{"image_id": 391895, "caption": captions[0]}, {"image_id": 522418, "caption": captions[1]}]
    # test: the csv pandas writes for DataFrame({'image_files', 'caption', 'prob'}) (base_model.py:157-160)
    q = str(tmp_path / "results.csv")
    files = ["./test/images/1.jpg", "./test/images/2.jpg"]
    sat_b200.write_test_results(q, files, captions, scores)
    ref = str(tmp_path / "ref.csv")
    pd.DataFrame({"image_files": files, "caption": captions, "prob": scores}).to_csv(ref)
    a, b = list(csv.reader(open(q))), list(csv.reader(open(ref)))
    assert a[0] == b[0] and len(a) == len(b)
    for ra, rb in zip(a[1:], b[1:]):
        assert ra[:3] == rb[:3] and float(ra[3]) == float(rb[3])


def test_learning_rate_staircase_schedule():
    """model.py:466-476: exponential_decay(lr, global_step, decay_steps, decay_rate, staircase=True)."""
    class M(object):
        learning_rate = sat_b200.CaptionGenerator.learning_rate
        global_step = 0
    m = M()
    m.config = sat_b200.Config()
    assert m.learning_rate(123456) == 1e-4                        # default factor 1.0: no decay
    m.config = sat_b200.Config(learning_rate_decay_factor=0.5, num_steps_per_decay=100)
    for step, want in ((0, 1e-4), (99, 1e-4), (100, 5e-5), (250, 2.5e-5)):
        assert abs(m.learning_rate(step) - want) < 1e-18
    assert m.config.apply_learning_rate_decay is False             # reference: the optimizer keeps the initial rate
