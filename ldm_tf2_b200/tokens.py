"""Tokenizer known-answer ids of the reference's default prompt and of the empty prompt
(convert_ckpt_pytorch_to_tf2.py:384-392): BERT-uncased WordPiece, max_length 77.  These constants let
benchmarks and tests run where the vocab file is not available; get_token_ids tokenises any prompt."""
import numpy as np

DEFAULT_PROMPT = "a virus monster is playing guitar, oil on canvas"
COND_IDS = [101, 1037, 7865, 6071, 2003, 2652, 2858, 1010, 3514, 2006, 10683, 102] + [0] * 65
UNCOND_IDS = [101, 102] + [0] * 75


def default_token_ids(batch_size: int) -> np.ndarray:
    """get_token_ids layout (run_ldm_sampler.py:42-45): B uncond rows then B cond rows, int64."""
    return np.array([UNCOND_IDS] * batch_size + [COND_IDS] * batch_size, dtype=np.int64)


def get_token_ids(prompt: str, vocab_dir: str, batch_size: int, max_length: int = 77) -> np.ndarray:
    """run_ldm_sampler.py:28-46 over the same vocab.txt, with the standalone WordPiece tokenizer
    (ldm_tf2_b200/wordpiece.py; pinned against HF's BertTokenizerFast by tests/golden/wordpiece_small.json)."""
    from . import wordpiece
    return wordpiece.get_token_ids(prompt, vocab_dir, batch_size, max_length)
