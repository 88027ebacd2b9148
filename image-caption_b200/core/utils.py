"""Host-side helpers kept from the reference's core/utils.py: decode_captions (utils.py:67-103),
pickle helpers (utils.py:106-117), write_scores (utils.py:120-138)."""
import os
import pickle

import numpy as np


def decode_captions(captions, index_to_word):
    """ids -> strings; skips a leading <START>, stops at <END> (emitted as '.'), drops <NULL>."""
    captions = np.asarray(captions)
    if captions.ndim == 1:
        captions = captions[None, :]
    decoded = []
    for row in captions:
        words = []
        for t, idx in enumerate(row):
            word = index_to_word[int(idx)]
            if word == '<START>' and t == 0:
                continue
            if word == '<END>':
                words.append('.')
                break
            if word != '<NULL>':
                words.append(word)
        decoded.append(' '.join(words))
    return decoded


def load_pickle(path):
    with open(path, 'rb') as f:
        return pickle.load(f)


def save_pickle(data, path):
    with open(path, 'wb') as f:
        pickle.dump(data, f, pickle.HIGHEST_PROTOCOL)


def write_scores(scores, path, epoch, split):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, f'{split}_scores.txt'), 'a') as f:
        f.write(f'Epoch {epoch}\n')
        for k, v in scores.items():
            f.write(f'{k}: {v}\n')
        f.write('\n')
