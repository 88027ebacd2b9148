"""Flat module constants, same names as the reference's core/config.py (config.py:5-129).  Only the values
the caption-model wrapper and the CLI read are kept; select the model block with ICAP_OUTPUT_NAME."""
import os

import torch

# preprocess (config.py:5-9)
MAX_LENGTH = int(os.environ.get("ICAP_MAX_LENGTH", 49))
WORD_COUNT_THRESHOLD = 1
NUM_OBJECT = 36
PAD_IDX = 0
MAX_OBJ = 5

IMAGE_MODEL = 'YOLOv5'
# 'Transformer' (teacher-forced CE) or 'RL_Transformer' (SelfCriticNetwork: PolicyNetwork + self-critical loss, the
# reference's shipped default, config.py:14).  The default here is the north_star's Transformer path.
CAPTION_MODEL = os.environ.get("ICAP_CAPTION_MODEL", 'Transformer')
assert CAPTION_MODEL in ('Transformer', 'RL_Transformer')

MODEL_NAME = 'maxlen49_36obj_1wordCount'
OUTPUT_NAME = os.environ.get("ICAP_OUTPUT_NAME", 'maxlen49_36obj_1wordCount_256_25b_32h_split_img_obj')

DATA_PATH = f'./data/{MODEL_NAME}'
OUTPUT_PATH = f'./output/{OUTPUT_NAME}'
WORD_TO_IDX_PATH = f'{DATA_PATH}/train/word_index.pkl'
SYNTHETIC_VOCAB = int(os.environ.get("ICAP_SYNTHETIC_VOCAB", 10000))   # used when word_index.pkl is absent

# the reference pins cuda:2 (config.py:35); one process drives one GPU here.  There is no CPU execution path: without a
# GPU the device is still "cuda:0" and the first kernel launch fails loudly.
DEVICE = torch.device(os.environ.get("ICAP_DEVICE", "cuda:0"))

# encoder (config.py:51-56)
ENCODE_DIM_FEATURES = 2048
ENCODE_DIM_POSITIONS = 84 if IMAGE_MODEL == 'YOLOv5' else 95

# solver (config.py:59-68)
NUM_EPOCH = int(os.environ.get("ICAP_NUM_EPOCH", 1000))
BATCH_SIZE = int(os.environ.get("ICAP_BATCH_SIZE", 32))
DROPOUT = 0.3
LEARNING_RATE = 0.0005
REGION_CACHE = os.environ.get("ICAP_REGION_CACHE", "1") != "0"    # keep each split's region features packed in HBM
LOG_PATH = f'./logs_{OUTPUT_NAME}/'
WRITE_LOG = ['loss']
if CAPTION_MODEL.find('RL') != -1:                 # config.py:65-68,81-85
    WRITE_LOG = ['loss', 'language_model_loss', 'structure_loss', 'reward']
    STRUCTURE_LOSS_WEIGHT = 0.5
    CIDER_REWARD_WEIGHT = 1
    BLEU_REWARD_WEIGHT = 1
    ENTROPY_REWARD_WEIGHT = 1
    SELF_CIDER_REWARD_WEIGHT = 1

if OUTPUT_NAME == 'maxlen49_36obj_1wordCount_256_25b_32h_split_img_obj':      # config.py:105-129
    MOVE_FIRST_IMAGE_FAETURE = False
    SPLIT_POSITION = False
    ENCODE_MASK = True
    SPLIT_IMAGE_OBJECTS = True
    ENCODE_INPUT_SIZE = 256
    ENCODE_Q_K_DIM = 256
    ENCODE_V_DIM = 256
    ENCODE_HIDDEN_SIZE = 256
    ENCODE_NUM_BLOCKS = 2
    ENCODE_NUM_HEADS = 32
    DIM_WORD_EMBEDDING = 256
    DECODE_INPUT_SIZE = 256
    DECODE_Q_K_DIM = 256
    DECODE_V_DIM = 256
    DECODE_HIDDEN_SIZE = 256
    DECODE_NUM_BLOCKS = 5
    DECODE_NUM_HEADS = 32
else:                                                                         # Transformer ctor defaults (model.py:15-36)
    MOVE_FIRST_IMAGE_FAETURE = False
    SPLIT_POSITION = False
    ENCODE_MASK = False
    SPLIT_IMAGE_OBJECTS = False
    ENCODE_INPUT_SIZE = 512
    ENCODE_Q_K_DIM = 512
    ENCODE_V_DIM = 512
    ENCODE_HIDDEN_SIZE = 2048
    ENCODE_NUM_BLOCKS = 6
    ENCODE_NUM_HEADS = 8
    DIM_WORD_EMBEDDING = 512
    DECODE_INPUT_SIZE = 512
    DECODE_Q_K_DIM = 512
    DECODE_V_DIM = 512
    DECODE_HIDDEN_SIZE = 2048
    DECODE_NUM_BLOCKS = 6
    DECODE_NUM_HEADS = 8
