"""SQuAD v1.1 / v2.0 data pipeline without third-party dependencies: JSON → examples → sliding-window features → n-best span
decoding → exact-match / F1.  The reference's example takes all of this from ``transformers`` (squad_convert_examples_to_features,
compute_predictions_logits, squad_evaluate — /root/reference/examples/squad/main.py:36-58,330-470); this file is an independent
implementation of the published BERT recipe (Devlin et al. 2018, §4.2; Rajpurkar et al. 2016/2018 for the metric) so the example
runs in an image that has neither a network nor a tokenizer download.

Tokenizers: :class:`WordPieceTokenizer` reads a BERT ``vocab.txt`` (greedy longest-match-first word pieces after lower-casing,
accent stripping and punctuation splitting); :class:`HashTokenizer` needs no file — words are hashed into a fixed id range — and
is what the offline tests and the synthetic data use."""
import collections
import json
import math
import re
import string
import unicodedata
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

# ---------------------------------------------------------------------------------------------------------------------
# tokenizers
# ---------------------------------------------------------------------------------------------------------------------


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(ch: str) -> bool:
    cp = ord(ch)
    return 0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or 0xF900 <= cp <= 0xFAFF


def basic_tokenize(text: str, lower: bool) -> List[str]:
    """Whitespace + punctuation split; control characters dropped, CJK characters isolated, accents stripped when lower-casing."""
    out, word = [], []

    def flush():
        if word:
            out.append("".join(word))
            word.clear()

    for ch in text:
        cat = unicodedata.category(ch)
        if ch in "\t\n\r " or cat == "Zs":
            flush()
        elif ord(ch) == 0 or ord(ch) == 0xFFFD or (cat.startswith("C") and ch not in "\t\n\r"):
            continue
        elif _is_punct(ch) or _is_cjk(ch):
            flush()
            out.append(ch)
        else:
            word.append(ch)
    flush()
    if lower:
        out = ["".join(c for c in unicodedata.normalize("NFD", w.lower()) if unicodedata.category(c) != "Mn") for w in out]
        out = [w for w in out if w]
    return out


class _Special:
    pad_token, unk_token, cls_token, sep_token = "[PAD]", "[UNK]", "[CLS]", "[SEP]"


class WordPieceTokenizer(_Special):
    """BERT word pieces from a ``vocab.txt`` (one token per line, line number = id)."""

    def __init__(self, vocab_file: str, do_lower_case: bool = True, max_chars_per_word: int = 100):
        with open(vocab_file, encoding="utf-8") as f:
            self.vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        self.lower, self.max_chars = do_lower_case, max_chars_per_word
        self.pad_id, self.unk_id, self.cls_id, self.sep_id = (self.vocab[t] for t in (self.pad_token, self.unk_token, self.cls_token, self.sep_token))

    @property
    def vocab_size(self) -> int:
        return len(self.vocab)

    def _pieces(self, word: str) -> List[str]:
        if len(word) > self.max_chars:
            return [self.unk_token]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end] if start == 0 else "##" + word[start:end]
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk_token]
            pieces.append(cur)
            start = end
        return pieces

    def tokenize(self, text: str) -> List[str]:
        return [p for w in basic_tokenize(text, self.lower) for p in self._pieces(w)]

    def convert_tokens_to_ids(self, tokens: Sequence[str]) -> List[int]:
        return [self.vocab.get(t, self.unk_id) for t in tokens]


class HashTokenizer(_Special):
    """File-less stand-in: a word is one token, its id a CRC of the (lower-cased) word folded into ``[4, vocab_size)``."""

    def __init__(self, vocab_size: int = 30522, do_lower_case: bool = True):
        assert vocab_size > 8
        self._n, self.lower = vocab_size, do_lower_case
        self.pad_id, self.unk_id, self.cls_id, self.sep_id = 0, 1, 2, 3

    @property
    def vocab_size(self) -> int:
        return self._n

    def tokenize(self, text: str) -> List[str]:
        return basic_tokenize(text, self.lower)

    def convert_tokens_to_ids(self, tokens: Sequence[str]) -> List[int]:
        special = {self.pad_token: 0, self.unk_token: 1, self.cls_token: 2, self.sep_token: 3}
        return [special[t] if t in special else 4 + zlib.crc32(t.encode("utf-8")) % (self._n - 4) for t in tokens]


def load_tokenizer(name_or_path: Optional[str], do_lower_case: bool, vocab_size: int):
    """``vocab.txt`` file or a directory holding one → word pieces; anything else → the hashing tokenizer."""
    import os

    if name_or_path:
        cand = name_or_path if os.path.isfile(name_or_path) else os.path.join(name_or_path, "vocab.txt")
        if os.path.isfile(cand):
            return WordPieceTokenizer(cand, do_lower_case)
    return HashTokenizer(vocab_size, do_lower_case)


# ---------------------------------------------------------------------------------------------------------------------
# examples
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class SquadExample:
    qas_id: str
    question: str
    doc_tokens: List[str]                 # context split on whitespace
    answers: List[str] = field(default_factory=list)   # all reference answers (evaluation)
    answer_text: Optional[str] = None     # the training answer
    start_word: int = -1
    end_word: int = -1
    is_impossible: bool = False


def _split_context(context: str) -> Tuple[List[str], List[int]]:
    """Whitespace tokens of the context and, per character, the index of the token it belongs to."""
    tokens, char_to_word, in_word = [], [], False
    for ch in context:
        if ch in " \t\r\n" or ord(ch) in (0x202F, 0xA0):
            in_word = False
        else:
            if not in_word:
                tokens.append(ch)
            else:
                tokens[-1] += ch
            in_word = True
        char_to_word.append(len(tokens) - 1)
    return tokens, char_to_word


def read_squad_examples(path_or_dict, is_training: bool, version_2: bool = False) -> List[SquadExample]:
    data = path_or_dict
    if isinstance(path_or_dict, str):
        with open(path_or_dict, encoding="utf-8") as f:
            data = json.load(f)
    out = []
    for article in data["data"]:
        for para in article["paragraphs"]:
            tokens, c2w = _split_context(para["context"])
            for qa in para["qas"]:
                ex = SquadExample(qas_id=str(qa["id"]), question=qa["question"], doc_tokens=tokens, is_impossible=bool(qa.get("is_impossible", False)) and version_2)
                ex.answers = [a["text"] for a in qa.get("answers", [])]
                if not ex.is_impossible and qa.get("answers"):
                    a = qa["answers"][0]
                    start_char = int(a["answer_start"])
                    end_char = min(start_char + len(a["text"]) - 1, len(c2w) - 1)
                    ex.answer_text, ex.start_word, ex.end_word = a["text"], max(c2w[start_char], 0), max(c2w[end_char], 0)
                    if is_training:   # drop answers that cannot be recovered from the whitespace tokens (mis-aligned annotations)
                        have = " ".join(tokens[ex.start_word: ex.end_word + 1])
                        want = " ".join(a["text"].split())
                        if want not in have:
                            continue
                elif is_training and not version_2:
                    continue
                out.append(ex)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# features
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class SquadFeature:
    unique_id: int
    example_index: int
    tokens: List[str]
    token_to_orig: Dict[int, int]          # position in the window → whitespace-token index of the context
    token_is_max_context: Dict[int, bool]
    input_ids: List[int]
    attention_mask: List[int]
    token_type_ids: List[int]
    start_position: int = 0
    end_position: int = 0
    is_impossible: bool = False


def _tighten_answer(sub_tokens, lo, hi, tokenizer, answer_text):
    """The annotated answer may be a part of a whitespace token ("1895." vs "1895"): shrink the sub-token span to the tightest one
    whose pieces spell the tokenised answer."""
    want = " ".join(tokenizer.tokenize(answer_text))
    for s in range(lo, hi + 1):
        for e in range(hi, s - 1, -1):
            if " ".join(sub_tokens[s: e + 1]) == want:
                return s, e
    return lo, hi


def convert_examples_to_features(examples: Sequence[SquadExample], tokenizer, max_seq_length: int = 384, doc_stride: int = 128, max_query_length: int = 64,
                                 is_training: bool = True, first_unique_id: int = 1_000_000_000) -> List[SquadFeature]:
    """``[CLS] question [SEP] context-window [SEP]`` windows of at most ``max_seq_length`` tokens that advance by ``doc_stride`` context
    tokens.  A window that does not contain the whole answer is labelled with position 0 (the [CLS] token), which is also the
    label of unanswerable questions."""
    feats, uid = [], first_unique_id
    for ex_i, ex in enumerate(examples):
        q_tokens = tokenizer.tokenize(ex.question)[:max_query_length]
        sub_tokens, sub_to_word, word_to_sub = [], [], []
        for w_i, w in enumerate(ex.doc_tokens):
            word_to_sub.append(len(sub_tokens))
            for piece in tokenizer.tokenize(w):
                sub_tokens.append(piece)
                sub_to_word.append(w_i)
        ans_lo = ans_hi = -1
        if is_training and not ex.is_impossible and ex.start_word >= 0:
            ans_lo = word_to_sub[ex.start_word]
            ans_hi = word_to_sub[ex.end_word + 1] - 1 if ex.end_word + 1 < len(ex.doc_tokens) else len(sub_tokens) - 1
            ans_lo, ans_hi = _tighten_answer(sub_tokens, ans_lo, max(ans_hi, ans_lo), tokenizer, ex.answer_text)
        room = max_seq_length - len(q_tokens) - 3
        assert room > 0, "max_seq_length too small for the question"
        spans, start = [], 0
        while True:
            length = min(room, len(sub_tokens) - start)
            spans.append((start, length))
            if start + length >= len(sub_tokens):
                break
            start += min(length, doc_stride)
        for span_i, (s0, ln) in enumerate(spans):
            tokens = [tokenizer.cls_token] + q_tokens + [tokenizer.sep_token]
            types = [0] * len(tokens)
            offset = len(tokens)
            t2o, tmax = {}, {}
            for k in range(ln):
                pos = s0 + k
                t2o[offset + k] = sub_to_word[pos]
                tmax[offset + k] = _is_max_context(spans, span_i, pos)
                tokens.append(sub_tokens[pos])
                types.append(1)
            tokens.append(tokenizer.sep_token)
            types.append(1)
            ids = tokenizer.convert_tokens_to_ids(tokens)
            mask = [1] * len(ids)
            pad = max_seq_length - len(ids)
            ids, mask, types = ids + [tokenizer.pad_id] * pad, mask + [0] * pad, types + [0] * pad
            sp = ep = 0
            impossible = ex.is_impossible
            if is_training and not ex.is_impossible:
                if ans_lo >= s0 and ans_hi <= s0 + ln - 1:
                    sp, ep = ans_lo - s0 + offset, ans_hi - s0 + offset
                else:
                    impossible = True   # this window misses the answer: target is [CLS]
            feats.append(SquadFeature(uid, ex_i, tokens, t2o, tmax, ids, mask, types, sp, ep, impossible))
            uid += 1
    return feats


def _is_max_context(spans, cur, pos) -> bool:
    """A context token appears in several overlapping windows; only the window where it has the most context on both sides may vote for it."""
    best, best_i = None, None
    for i, (s0, ln) in enumerate(spans):
        end = s0 + ln - 1
        if pos < s0 or pos > end:
            continue
        score = min(pos - s0, end - pos) + 0.01 * ln
        if best is None or score > best:
            best, best_i = score, i
    return best_i == cur


def features_to_tensors(feats: Sequence[SquadFeature], is_training: bool) -> Dict[str, torch.Tensor]:
    t = {"input_ids": torch.tensor([f.input_ids for f in feats], dtype=torch.long), "token_type_ids": torch.tensor([f.token_type_ids for f in feats], dtype=torch.long),
         "attention_mask": torch.tensor([f.attention_mask for f in feats], dtype=torch.long), "feature_index": torch.arange(len(feats), dtype=torch.long)}
    if is_training:
        t["start_positions"] = torch.tensor([f.start_position for f in feats], dtype=torch.long)
        t["end_positions"] = torch.tensor([f.end_position for f in feats], dtype=torch.long)
    return t


# ---------------------------------------------------------------------------------------------------------------------
# decoding
# ---------------------------------------------------------------------------------------------------------------------
def _top(logits: Sequence[float], n: int) -> List[int]:
    return sorted(range(len(logits)), key=lambda i: logits[i], reverse=True)[:n]


def _edge_trim(f: "SquadFeature", s: int, e: int) -> Tuple[int, int]:
    """A span may begin / end inside a whitespace token ("(1895)," when the model points at "1895"): number of leading / trailing
    characters of the first / last word to drop — only punctuation pieces that lie outside the span are ever dropped."""
    front = back = 0
    k = s - 1
    while k in f.token_to_orig and f.token_to_orig[k] == f.token_to_orig[s] and len(f.tokens[k]) == 1 and _is_punct(f.tokens[k]):
        front += 1
        k -= 1
    if k in f.token_to_orig and f.token_to_orig[k] == f.token_to_orig[s]:
        front = 0   # a non-punctuation piece of the same word precedes the span: keep the word whole
    k = e + 1
    while k in f.token_to_orig and f.token_to_orig[k] == f.token_to_orig[e] and len(f.tokens[k]) == 1 and _is_punct(f.tokens[k]):
        back += 1
        k += 1
    if k in f.token_to_orig and f.token_to_orig[k] == f.token_to_orig[e]:
        back = 0
    return front, back


def compute_predictions(examples: Sequence[SquadExample], feats: Sequence[SquadFeature], logits: Dict[int, Tuple[Sequence[float], Sequence[float]]],
                        n_best_size: int = 20, max_answer_length: int = 30, version_2: bool = False, null_score_diff_threshold: float = 0.0,
                        return_nbest: bool = False):
    """``logits[unique_id] = (start_logits, end_logits)`` → ``{qas_id: answer text}``.  Per example: the ``n_best_size`` start and end
    positions of every window are paired, pairs outside the context / reversed / too long / not at their max-context window are
    dropped, the best remaining pair (summed logit) gives the answer as the whitespace tokens it covers.  With ``version_2`` the
    [CLS]+[CLS] score is the "no answer" candidate and wins when it beats the best span by more than the threshold."""
    by_example = collections.defaultdict(list)
    for f in feats:
        by_example[f.example_index].append(f)
    preds, nbest_out = {}, {}
    for ex_i, ex in enumerate(examples):
        cands, null_score = [], math.inf
        for f in by_example.get(ex_i, []):
            if f.unique_id not in logits:
                continue
            s_log, e_log = logits[f.unique_id]
            if version_2:
                null_score = min(null_score, s_log[0] + e_log[0])
            for s in _top(s_log, n_best_size):
                for e in _top(e_log, n_best_size):
                    if s not in f.token_to_orig or e not in f.token_to_orig or not f.token_is_max_context.get(s, False) or e < s or e - s + 1 > max_answer_length:
                        continue
                    cands.append((s_log[s] + e_log[e], f.token_to_orig[s], f.token_to_orig[e], *_edge_trim(f, s, e)))
        cands.sort(key=lambda c: c[0], reverse=True)
        seen, nbest = set(), []
        for score, ws, we, cut_front, cut_back in cands:
            words = list(ex.doc_tokens[ws: we + 1])
            if cut_front and len(words[0]) > cut_front:
                words[0] = words[0][cut_front:]
            if cut_back and len(words[-1]) > cut_back:
                words[-1] = words[-1][: len(words[-1]) - cut_back]
            text = " ".join(words)
            if text in seen:
                continue
            seen.add(text)
            nbest.append((text, score))
            if len(nbest) >= n_best_size:
                break
        if not nbest:
            nbest = [("", 0.0)]
        best_text, best_score = nbest[0]
        if version_2 and null_score - best_score > null_score_diff_threshold:
            best_text = ""
        preds[ex.qas_id] = best_text
        if return_nbest:
            z = max(s for _, s in nbest)
            den = sum(math.exp(s - z) for _, s in nbest)
            nbest_out[ex.qas_id] = [{"text": t, "probability": math.exp(s - z) / den, "logit": s} for t, s in nbest]
    return (preds, nbest_out) if return_nbest else preds


# ---------------------------------------------------------------------------------------------------------------------
# metric
# ---------------------------------------------------------------------------------------------------------------------
def normalize_answer(s: str) -> str:
    s = "".join(ch for ch in s.lower() if ch not in set(string.punctuation))
    s = re.sub(r"\b(a|an|the)\b", " ", s)
    return " ".join(s.split())


def _f1(pred: str, truth: str) -> float:
    p, t = normalize_answer(pred).split(), normalize_answer(truth).split()
    if not p or not t:
        return float(p == t)
    common = collections.Counter(p) & collections.Counter(t)
    same = sum(common.values())
    if same == 0:
        return 0.0
    precision, recall = same / len(p), same / len(t)
    return 2 * precision * recall / (precision + recall)


def squad_evaluate(examples: Sequence[SquadExample], preds: Dict[str, str]) -> Dict[str, float]:
    """Exact match and token F1 in percent, the maximum over the reference answers of each question (unanswerable: the empty string)."""
    em = f1 = 0.0
    n = 0
    for ex in examples:
        if ex.qas_id not in preds:
            continue
        truths = [a for a in ex.answers if normalize_answer(a)] or [""]
        if ex.is_impossible:
            truths = [""]
        p = preds[ex.qas_id]
        em += max(float(normalize_answer(p) == normalize_answer(t)) for t in truths)
        f1 += max(_f1(p, t) for t in truths)
        n += 1
    return {"exact": 100.0 * em / max(n, 1), "f1": 100.0 * f1 / max(n, 1), "total": n}


# ---------------------------------------------------------------------------------------------------------------------
# offline data
# ---------------------------------------------------------------------------------------------------------------------
_WORDS = ("river mountain castle engine garden library market harbor bridge forest village desert island valley tower meadow canyon glacier "
          "orchard lantern compass anchor violin hammer ladder mirror basket candle feather pebble ribbon saddle thimble walnut zephyr quartz").split()
_NAMES = "Avery Blake Casey Devon Ellis Finley Harper Jordan Kendall Logan Morgan Parker Quinn Reese Sawyer Taylor".split()
_COLOURS = "red blue green amber violet silver golden crimson ivory teal".split()


def synthetic_squad(n_paragraphs: int, seed: int = 0, questions_per_paragraph: int = 3, version_2: bool = False) -> dict:
    """A SQuAD-format dictionary generated from templates: every paragraph states facts ("The castle of Avery is crimson.") among
    filler sentences and asks for them ("What colour is the castle of Avery?"), so a model can learn the task and the whole
    pipeline — alignment, windows, decoding, metric — is exercised with real text."""
    import random

    rng = random.Random(seed)
    data = []
    for p_i in range(n_paragraphs):
        facts, sentences = [], []
        for _ in range(questions_per_paragraph + 2):
            thing, name, colour, count = rng.choice(_WORDS), rng.choice(_NAMES), rng.choice(_COLOURS), rng.randint(2, 97)
            kind = rng.randrange(3)
            if kind == 0:
                sentences.append(f"The {thing} of {name} is {colour}.")
                facts.append((f"What colour is the {thing} of {name}?", colour))
            elif kind == 1:
                sentences.append(f"{name} keeps {count} {thing}s near the {rng.choice(_WORDS)}.")
                facts.append((f"How many {thing}s does {name} keep?", str(count)))
            else:
                place = rng.choice(_WORDS)
                sentences.append(f"Every spring {name} walks from the {thing} to the old {place}.")
                facts.append((f"Where does {name} walk to from the {thing}?", f"the old {place}"))
            sentences.append(" ".join(rng.choice(_WORDS) for _ in range(rng.randint(4, 12))).capitalize() + ".")
        context = " ".join(sentences)
        qas = []
        for q_i, (question, answer) in enumerate(facts[:questions_per_paragraph]):
            start = context.find(answer)
            qas.append({"id": f"syn-{seed}-{p_i}-{q_i}", "question": question, "answers": [{"text": answer, "answer_start": start}], "is_impossible": False})
        if version_2:
            qas.append({"id": f"syn-{seed}-{p_i}-na", "question": f"Who painted the {rng.choice(_WORDS)} of nobody?", "answers": [], "is_impossible": True})
        data.append({"title": f"synthetic {p_i}", "paragraphs": [{"context": context, "qas": qas}]})
    return {"version": "v2.0" if version_2 else "1.1", "data": data}
