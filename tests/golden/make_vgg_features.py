"""Real conv5_3 features and real caption ids from the data the reference ships.

Run HERE (the build container), never on the GPU box:
    python tests/golden/make_vgg_features.py
Reads (read-only) /root/reference/data/vgg16_no_fc.npy (13 conv kernels [3,3,Cin,Cout] + biases),
utils/ilsvrc_2012_mean.npy, the 11 + 3 images under data/val/images and data/test/images, data/vocabulary.csv and
data/train/captions_train2014.json, and writes tests/golden/vgg_conv5_3.npz:

  feats      float16 [14, 196, 512]  conv5_3 activations = the decoder's `contexts` (model.py:32-59)
  sentences  int32   [40, 20]        word ids of 40 real COCO train captions (dataset.py:133-146), 0 padded
  masks      float32 [40, 20]
  stats      json string: mean / sparsity / quantiles of the features (used to shape VGG-like synthetic contexts)

The CNN is OUT OF SCOPE of the build; this torch-CPU restatement of model.py:32-59 (conv2d 3x3 'same' + ReLU, 2x2 max
pools) exists only to give the oracle REAL feature statistics: the one training step the reference recorded
(summary/events.out.tfevents..., B=20: cross_entropy 8.7459, gradient_norm 2.5466, attention_loss 0.004455) was taken on
real conv5_3 features, and tests/test_oracle_recorded_step.py brackets those numbers with the training oracle fed
from this file.  Image preprocessing follows utils/misc.py:6-28 (channel flip, 224x224 bilinear resize, mean file
subtraction) with torch's bilinear (half-pixel centres, no antialias = cv2.INTER_LINEAR's convention); captions are
tokenised with a regex stand-in for nltk.word_tokenize (not installed) and captions with unknown words or more than
20 tokens are skipped (dataset.py filter_by_cap_len / filter_by_words).
"""
import json
import os
import re

import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vgg_conv5_3.npz")
ORDER = ["conv1_1", "conv1_2", "P", "conv2_1", "conv2_2", "P", "conv3_1", "conv3_2", "conv3_3", "P",
         "conv4_1", "conv4_2", "conv4_3", "P", "conv5_1", "conv5_2", "conv5_3"]


def load_image(path, mean):
    img = np.asarray(Image.open(path).convert("RGB"), np.float32)          # cv2.imread + the channel flip = RGB
    x = torch.from_numpy(img).permute(2, 0, 1)[None]
    x = F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False)
    return x - torch.from_numpy(mean.astype(np.float32))[None, :, None, None]   # (the reference subtracts as is)


def main():
    torch.set_grad_enabled(False)
    w = np.load(os.path.join(REF, "data/vgg16_no_fc.npy"), encoding="latin1", allow_pickle=True).item()
    mean = np.load(os.path.join(REF, "utils/ilsvrc_2012_mean.npy")).mean(1).mean(1)
    files = sorted(os.path.join(REF, "data/val/images", f) for f in os.listdir(os.path.join(REF, "data/val/images"))
                   if f.endswith(".jpg"))
    files += [os.path.join(REF, "data/test/images", "%d.jpg" % i) for i in (1, 2, 3)]
    feats = []
    for f in files:
        x = load_image(f, mean)
        for name in ORDER:
            if name == "P":
                x = F.max_pool2d(x, 2, 2)
            else:
                k = torch.from_numpy(np.ascontiguousarray(w[name]["kernel"].transpose(3, 2, 0, 1)).astype(np.float32))
                x = F.relu(F.conv2d(x, k, torch.from_numpy(w[name]["bias"].astype(np.float32)), padding=1))
        feats.append(x[0].permute(1, 2, 0).reshape(196, 512).numpy())     # NHWC [14,14,512] -> [196,512] (model.py:54-55)
        print(os.path.basename(f), "mean %.3f  zeros %.3f  max %.1f" % (feats[-1].mean(), (feats[-1] == 0).mean(), feats[-1].max()))
    feats = np.stack(feats)

    import csv
    word2idx = {}
    with open(os.path.join(REF, "data/vocabulary.csv")) as fh:
        for row in csv.DictReader(fh):
            word2idx[row["word"]] = int(row["index"])
    ann = json.load(open(os.path.join(REF, "data/train/captions_train2014.json")))["annotations"]
    sents, masks = [], []
    for a in ann:
        toks = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", a["caption"].lower())
        if not toks or len(toks) > 20 or any(t not in word2idx for t in toks):
            continue
        ids = np.zeros(20, np.int32)
        ids[:len(toks)] = [word2idx[t] for t in toks]
        mk = np.zeros(20, np.float32)
        mk[:len(toks)] = 1.0
        sents.append(ids)
        masks.append(mk)
        if len(sents) == 40:
            break
    q = [50, 90, 99, 99.9]
    stats = dict(mean=float(feats.mean()), std=float(feats.std()), zero_fraction=float((feats == 0).mean()),
                 max=float(feats.max()), quantiles={str(p): float(np.percentile(feats, p)) for p in q},
                 mean_caption_length=float(np.mean([m.sum() for m in masks])), images=[os.path.basename(f) for f in files])
    np.savez_compressed(OUT, feats=feats.astype(np.float16), sentences=np.stack(sents), masks=np.stack(masks),
                        stats=json.dumps(stats))
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB", stats)


if __name__ == "__main__":
    main()
