"""Host-side helpers kept from the reference's core/utils.py: load_coco_data (utils.py:32-64), decode_captions
(utils.py:67-103), pickle helpers (utils.py:106-117), write_scores (utils.py:120-138)."""
import os
import pickle
import time

import numpy as np


def _load_array(stem):
    """`<stem>.hkl` through hickle (the reference's container, utils.py:46-47) when hickle is installed; otherwise
    `<stem>.npy`, memory-mapped so a 30 GB feature file is streamed into the device cache without a host copy."""
    if os.path.exists(stem + '.npy'):
        return np.load(stem + '.npy', mmap_mode='r')
    if os.path.exists(stem + '.hkl'):
        try:
            import hickle
        except ImportError as exc:
            raise ImportError(f'{stem}.hkl needs the `hickle` package (not installable offline); convert it once with '
                              f'`np.save("{stem}.npy", hickle.load("{stem}.hkl"))`') from exc
        return hickle.load(stem + '.hkl')
    raise FileNotFoundError(f'{stem}.hkl / {stem}.npy not found')


def load_coco_data(data_path='./data/MSCOCO', split='train'):
    """Same dict as the reference: features [n_img, R, 2048] f32, positions [n_img, R, Dp] f32, file_names, captions
    [n_cap, max_length+2] int32, image_idxs [n_cap] (+ word_to_idx for the train split)."""
    data_path = os.path.join(data_path, split)
    start_time = time.time()
    data = {'features': _load_array(os.path.join(data_path, f'{split}.features')),
            'positions': _load_array(os.path.join(data_path, f'{split}.positions'))}
    for key, name in (('file_names', 'file.names'), ('captions', 'captions'), ('image_idxs', 'image.indices')):
        data[key] = load_pickle(os.path.join(data_path, f'{split}.{name}.pkl'))
    if split == 'train':
        data['word_to_idx'] = load_pickle(os.path.join(data_path, 'word_index.pkl'))
    for key, value in data.items():
        if isinstance(value, np.ndarray):
            print(key, type(value), value.shape, value.dtype)
        else:
            print(key, type(value), len(value))
    print('Elapse time: %.2f' % (time.time() - start_time))
    return data


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
