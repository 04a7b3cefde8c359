"""examples/squad/squad_utils.py: tokenizers, answer alignment, sliding windows, n-best decoding and the SQuAD metric — on text, with
oracle logits (a "model" that puts all its mass on the labelled positions must score 100)."""
import importlib.util
import os

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("squad_utils", os.path.join(REPO, "examples", "squad", "squad_utils.py"))
su = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(su)

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "the", "castle", "of", "avery", "is", "crimson", ".", ",", "(", ")", "what", "colour", "?", "un", "##aff", "##able",
         "built", "in", "1895", "was", "it", "when", "river", "##s", "walk", "##ing"]


@pytest.fixture
def wordpiece(tmp_path):
    path = tmp_path / "vocab.txt"
    path.write_text("\n".join(VOCAB) + "\n", encoding="utf-8")
    return su.WordPieceTokenizer(str(path), do_lower_case=True)


def test_wordpiece_is_greedy_longest_match_and_handles_accents_punctuation_unknowns(wordpiece):
    assert wordpiece.tokenize("Unaffable rivers, walking!") == ["un", "##aff", "##able", "river", "##s", ",", "walk", "##ing", "[UNK]"]
    assert wordpiece.tokenize("Thé castle (1895).") == ["the", "castle", "(", "1895", ")", "."]
    assert wordpiece.tokenize("zzz") == ["[UNK]"] and wordpiece.tokenize("  \t\n") == []
    ids = wordpiece.convert_tokens_to_ids(["[CLS]", "castle", "nonsense", "[SEP]"])
    assert ids == [2, 5, 1, 3] and wordpiece.vocab_size == len(VOCAB)
    assert isinstance(su.load_tokenizer(None, True, 100), su.HashTokenizer) and isinstance(su.load_tokenizer("bert-large-uncased", True, 100), su.HashTokenizer)


def test_hash_tokenizer_is_deterministic_and_keeps_special_ids():
    a, b = su.HashTokenizer(1000), su.HashTokenizer(1000)
    toks = a.tokenize("The Castle of Avery, 1895!")
    assert toks == ["the", "castle", "of", "avery", ",", "1895", "!"]
    ids = a.convert_tokens_to_ids(["[CLS]"] + toks + ["[SEP]", "[PAD]"])
    assert ids == b.convert_tokens_to_ids(["[CLS]"] + toks + ["[SEP]", "[PAD]"])
    assert ids[0] == 2 and ids[-2] == 3 and ids[-1] == 0 and all(4 <= i < 1000 for i in ids[1:-2])


def _tiny_dataset():
    context = "The castle of Avery is crimson. It was built in (1895), when the river was low."
    return {"version": "1.1", "data": [{"title": "t", "paragraphs": [{"context": context, "qas": [
        {"id": "q1", "question": "What colour is the castle?", "answers": [{"text": "crimson", "answer_start": context.index("crimson")}]},
        {"id": "q2", "question": "When was it built?", "answers": [{"text": "1895", "answer_start": context.index("1895")}, {"text": "in 1895", "answer_start": 0}]},
    ]}]}]}


def test_answers_align_to_sub_tokens_inside_punctuated_words(wordpiece):
    exs = su.read_squad_examples(_tiny_dataset(), is_training=True)
    assert [e.qas_id for e in exs] == ["q1", "q2"] and exs[0].doc_tokens[5] == "crimson." and exs[1].doc_tokens[exs[1].start_word] == "(1895),"
    feats = su.convert_examples_to_features(exs, wordpiece, max_seq_length=48, doc_stride=16, max_query_length=16, is_training=True)
    assert len(feats) == 2
    f1, f2 = feats
    assert f1.tokens[f1.start_position: f1.end_position + 1] == ["crimson"]
    assert f2.tokens[f2.start_position: f2.end_position + 1] == ["1895"]          # not "(", "1895", ")", ","
    assert len(f1.input_ids) == 48 and sum(f1.attention_mask) == len(f1.tokens) and f1.token_type_ids[:len(f1.tokens)].count(1) == len(f1.tokens) - 8
    # decoding the labelled positions returns the clean answer: the surrounding punctuation of "(1895)," is trimmed
    logits = {}
    for f in feats:
        s, e = [-10.0] * 48, [-10.0] * 48
        s[f.start_position], e[f.end_position] = 10.0, 10.0
        logits[f.unique_id] = (s, e)
    preds = su.compute_predictions(exs, feats, logits)
    assert preds == {"q1": "crimson", "q2": "1895"}
    assert su.squad_evaluate(exs, preds) == {"exact": 100.0, "f1": 100.0, "total": 2}


def test_sliding_windows_cover_the_context_and_label_only_windows_with_the_answer():
    tok = su.HashTokenizer(5000)
    data = su.synthetic_squad(6, seed=3)
    exs = su.read_squad_examples(data, is_training=True)
    feats = su.convert_examples_to_features(exs, tok, max_seq_length=40, doc_stride=12, max_query_length=12, is_training=True)
    by_ex = {}
    for f in feats:
        by_ex.setdefault(f.example_index, []).append(f)
    assert len(feats) > len(exs)                       # windows were needed
    for ex_i, ex in enumerate(exs):
        fs = by_ex[ex_i]
        covered = set()
        for f in fs:
            covered.update(f.token_to_orig.values())
        assert covered == set(range(len(ex.doc_tokens)))          # every context word is in some window
        labelled = [f for f in fs if not f.is_impossible]
        assert labelled, "the answer must be inside at least one window"
        for f in labelled:
            words = [ex.doc_tokens[f.token_to_orig[p]] for p in range(f.start_position, f.end_position + 1)]
            assert su.normalize_answer(ex.answer_text) in su.normalize_answer(" ".join(words))
        for f in fs:
            if f.is_impossible:
                assert f.start_position == 0 and f.end_position == 0
    # each context sub-token has exactly one max-context window (positions are window-relative; consecutive windows advance by doc_stride)
    for ex_i in by_ex:
        count = {}
        starts = [i * 12 for i in range(len(by_ex[ex_i]))]
        for f, s0 in zip(by_ex[ex_i], starts):
            first = min(f.token_to_orig)
            for pos, is_max in f.token_is_max_context.items():
                if is_max:
                    count[s0 + pos - first] = count.get(s0 + pos - first, 0) + 1
        assert count and set(count.values()) == {1}


def test_oracle_logits_score_100_and_noise_scores_less_with_and_without_unanswerable_questions():
    tok = su.HashTokenizer(5000)
    for v2 in (False, True):
        data = su.synthetic_squad(8, seed=11, version_2=v2)
        train = su.read_squad_examples(data, is_training=True, version_2=v2)
        feats = su.convert_examples_to_features(train, tok, max_seq_length=48, doc_stride=16, max_query_length=12, is_training=True)
        logits, g = {}, torch.Generator().manual_seed(0)
        for f in feats:
            s, e = torch.full((48,), -8.0), torch.full((48,), -8.0)
            s[f.start_position], e[f.end_position] = 8.0, 8.0
            logits[f.unique_id] = (s.tolist(), e.tolist())
        preds, nbest = su.compute_predictions(train, feats, logits, version_2=v2, return_nbest=True)
        score = su.squad_evaluate(train, preds)
        assert score["exact"] == 100.0 and score["f1"] == 100.0 and score["total"] == len(train), (v2, score)
        assert all(abs(sum(c["probability"] for c in v) - 1.0) < 1e-6 for v in nbest.values())
        if v2:
            assert any(e.is_impossible for e in train) and all(preds[e.qas_id] == "" for e in train if e.is_impossible)
        noise = {f.unique_id: (torch.randn(48, generator=g).tolist(), torch.randn(48, generator=g).tolist()) for f in feats}
        assert su.squad_evaluate(train, su.compute_predictions(train, feats, noise, version_2=v2))["f1"] < 60.0


def test_metric_normalisation_and_multiple_references():
    exs = [su.SquadExample("a", "q", ["x"], answers=["The Eiffel Tower", "Eiffel Tower, Paris"]), su.SquadExample("b", "q", ["x"], answers=[], is_impossible=True)]
    assert su.squad_evaluate(exs, {"a": "eiffel tower!", "b": ""}) == {"exact": 100.0, "f1": 100.0, "total": 2}
    half = su.squad_evaluate(exs, {"a": "tower of Paris", "b": "something"})
    assert half["exact"] == 0.0 and 0.0 < half["f1"] < 50.0
    assert su.normalize_answer(" An  apple, the PIE. ") == "apple pie"


def test_features_to_tensors_shapes():
    tok = su.HashTokenizer(5000)
    exs = su.read_squad_examples(su.synthetic_squad(3, seed=1), is_training=True)
    feats = su.convert_examples_to_features(exs, tok, max_seq_length=64, doc_stride=32, max_query_length=16, is_training=True)
    t = su.features_to_tensors(feats, is_training=True)
    assert t["input_ids"].shape == (len(feats), 64) and t["start_positions"].dtype == torch.long and t["feature_index"].tolist() == list(range(len(feats)))
    assert (t["attention_mask"].sum(1) > 10).all() and ((t["token_type_ids"] == 1).sum(1) > 0).all()


def test_a_tiny_bert_learns_the_generated_dataset_through_the_whole_pipeline():
    """JSON → features → BertForQuestionAnswering → n-best decoding → metric: 200 AdamW steps take exact match on the training questions
    from ≈ 0 to well above 40 (3 s on a CPU) — alignment, labels, decoding and metric agree with each other."""
    from bagua_b200 import models

    torch.manual_seed(0)
    tok = su.HashTokenizer(1024)
    exs = su.read_squad_examples(su.synthetic_squad(24, seed=5), is_training=True)
    feats = su.convert_examples_to_features(exs, tok, max_seq_length=96, doc_stride=48, max_query_length=16)
    t = su.features_to_tensors(feats, True)
    model = models.BertForQuestionAnswering(models.BertConfig(vocab_size=1024, hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                                                              max_position_embeddings=96, hidden_dropout_prob=0.0))
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3)

    def score():
        model.eval()
        with torch.no_grad():
            s, e = model(t["input_ids"], token_type_ids=t["token_type_ids"], attention_mask=t["attention_mask"])
        model.train()
        return su.squad_evaluate(exs, su.compute_predictions(exs, feats, {f.unique_id: (s[i].tolist(), e[i].tolist()) for i, f in enumerate(feats)}))

    before = score()
    g = torch.Generator().manual_seed(1)
    for _ in range(200):
        idx = torch.randint(0, len(feats), (16,), generator=g)
        loss = model(t["input_ids"][idx], token_type_ids=t["token_type_ids"][idx], attention_mask=t["attention_mask"][idx],
                     start_positions=t["start_positions"][idx], end_positions=t["end_positions"][idx])[0]
        opt.zero_grad()
        loss.backward()
        opt.step()
    after = score()
    assert before["exact"] < 10.0 and after["exact"] > 40.0 and after["f1"] > after["exact"] - 1e-9, (before, after)
