"""Caption assembly after beam search: word ids -> sentence, and the result files of the reference's
eval / test loops.

  Vocabulary.load / get_sentence    utils/vocabulary.py:53-63, 72-80 (the csv written by Vocabulary.save)
  assemble_captions                 the per-image part of base_model.py:82-92 / 135-143: best caption + its score
  write_eval_results                base_model.py:109-111: json list of {"image_id", "caption"} (COCO result format)
  write_test_results                base_model.py:157-160: csv with columns image_files, caption, prob

Only the text side is mirrored: the reference also renders every image with its caption through matplotlib
(base_model.py:94-107, 145-155); image files are out of scope here (features are precomputed).  `beam_results` is what
CaptionGenerator.beam_search returns: per image, the captions sorted by descending score.
"""
import csv
import json
import string


class Vocabulary(object):
    """The word table of the reference (utils/vocabulary.py): `words[i]` is the word of id i; id 0 is '<start>' and
    '.' ends a sentence (id 2 in the shipped data/vocabulary.csv)."""

    def __init__(self, size=None, save_file=None, words=None):
        self.words = list(words) if words is not None else []
        self.word2idx = {w: i for i, w in enumerate(self.words)}
        self.size = size if size is not None else (len(self.words) or None)
        if save_file is not None:
            self.load(save_file)

    def load(self, save_file):
        """utils/vocabulary.py:72-80: the csv has the columns (index), frequency, index, word."""
        with open(save_file, newline="") as f:
            rows = list(csv.DictReader(f))
        rows.sort(key=lambda r: int(r["index"]))
        self.words = [r["word"] for r in rows]
        self.word2idx = {w: i for i, w in enumerate(self.words)}
        if self.size is None or self.size > len(self.words):
            self.size = len(self.words)
        return self

    @property
    def eos_id(self):
        """id of '.', the word that completes a caption in beam search (base_model.py:229)"""
        return self.word2idx["."]

    def get_sentence(self, idxs):
        """utils/vocabulary.py:53-63: words up to and including the first '.', a '.' appended if the last word is not
        one, joined with spaces except before punctuation and before tokens that start with an apostrophe."""
        words = [self.words[i] for i in idxs]
        if not words or words[-1] != ".":
            words.append(".")
        length = words.index(".") + 1
        words = words[:length]
        return "".join(" " + w if not w.startswith("'") and w not in string.punctuation else w for w in words).strip()


def assemble_captions(beam_results, vocabulary, fake_count=0):
    """Best caption and its score for every real image of a batch (base_model.py:82-92; the last batch of the reference
    is padded with `fake_count` copies, dataset.py:51-54, which are dropped here the same way)."""
    n = len(beam_results) - int(fake_count)
    captions, scores = [], []
    for caps in beam_results[:n]:
        best = caps[0]                                   # sorted by descending score (base_model.py:236-238)
        captions.append(vocabulary.get_sentence(best.sentence))
        scores.append(best.score)
    return captions, scores


def write_eval_results(path, image_ids, captions):
    """base_model.py:87-88, 109-111: [{"image_id": id, "caption": text}, ...] as json (what COCO.loadRes reads)."""
    results = [{"image_id": int(i), "caption": c} for i, c in zip(image_ids, captions)]
    with open(path, "w") as f:
        json.dump(results, f)
    return results


def write_test_results(path, image_files, captions, scores):
    """base_model.py:157-160: pandas.DataFrame({'image_files', 'caption', 'prob'}).to_csv(path) — a leading unnamed
    index column, then the columns in the order pandas keeps them (insertion order)."""
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["", "image_files", "caption", "prob"])
        for i, (a, c, p) in enumerate(zip(image_files, captions, scores)):
            w.writerow([i, a, c, repr(float(p))])
    return path
