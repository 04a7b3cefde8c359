"""bagua_b200.models.bert against HuggingFace's BertForQuestionAnswering (the model class of the reference's SQuAD example,
/root/reference/examples/squad/main.py:36-58 via ``transformers``): same weights → same logits and loss, in both directions of
the checkpoint conversion.  Random tiny weights: there is no network for pretrained ones."""
import json
import os

import pytest
import torch

transformers = pytest.importorskip("transformers")

from bagua_b200 import models  # noqa: E402


def _hf_model(layers=2, hidden=32, heads=4):
    cfg = transformers.BertConfig(vocab_size=120, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads, intermediate_size=2 * hidden,
                                  max_position_embeddings=48, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(3)
    return transformers.BertForQuestionAnswering(cfg).eval()


def _batch():
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(1, 120, (4, 24), generator=g)
    tt = (torch.arange(24) > 9).long().expand(4, 24).contiguous()
    am = torch.ones(4, 24, dtype=torch.long)
    am[1, 17:] = 0
    am[3, 20:] = 0
    sp, ep = torch.tensor([11, 12, 10, 15]), torch.tensor([13, 12, 16, 15])
    return ids, tt, am, sp, ep


def test_logits_and_loss_match_transformers(tmp_path):
    hm = _hf_model()
    hm.save_pretrained(str(tmp_path / "hf"))
    model, report = models.bert_qa_from_pretrained(str(tmp_path / "hf"))
    assert report == {"missing": [], "unexpected": []}
    model.eval()
    ids, tt, am, sp, ep = _batch()
    with torch.no_grad():
        want = hm(input_ids=ids, token_type_ids=tt, attention_mask=am, start_positions=sp, end_positions=ep)
        loss, start, end = model(ids, token_type_ids=tt, attention_mask=am, start_positions=sp, end_positions=ep)
    keep = am.bool()   # logits at padded positions are don't-cares (HF adds -inf style masks inside attention only)
    assert torch.allclose(start[keep], want.start_logits[keep], atol=1e-5) and torch.allclose(end[keep], want.end_logits[keep], atol=1e-5)
    assert abs(loss.item() - want.loss.item()) < 1e-5


def test_gradients_match_transformers(tmp_path):
    hm = _hf_model(layers=1).train()
    hm.save_pretrained(str(tmp_path / "hf"))
    model, _ = models.bert_qa_from_pretrained(str(tmp_path / "hf"))
    model.train()
    ids, tt, am, sp, ep = _batch()
    hm(input_ids=ids, token_type_ids=tt, attention_mask=am, start_positions=sp, end_positions=ep).loss.backward()
    model(ids, token_type_ids=tt, attention_mask=am, start_positions=sp, end_positions=ep)[0].backward()
    hq = torch.cat([getattr(hm.bert.encoder.layer[0].attention.self, n).weight.grad for n in ("query", "key", "value")])
    assert torch.allclose(model.layers[0].qkv.weight.grad, hq, atol=1e-6)
    assert torch.allclose(model.word.weight.grad, hm.bert.embeddings.word_embeddings.weight.grad, atol=1e-6)
    assert torch.allclose(model.qa_outputs.weight.grad, hm.qa_outputs.weight.grad, atol=1e-6)


def test_checkpoint_written_here_loads_in_transformers(tmp_path):
    torch.manual_seed(1)
    cfg = models.BertConfig(vocab_size=120, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, intermediate_size=64, max_position_embeddings=48, hidden_dropout_prob=0.0)
    model = models.BertForQuestionAnswering(cfg).eval()
    models.save_pretrained(model, str(tmp_path / "out"))
    with open(tmp_path / "out" / "config.json") as f:
        assert json.load(f)["model_type"] == "bert"
    hm = transformers.BertForQuestionAnswering.from_pretrained(str(tmp_path / "out")).eval()
    ids, tt, am, _, _ = _batch()
    with torch.no_grad():
        start, end = model(ids, token_type_ids=tt, attention_mask=am)
        want = hm(input_ids=ids, token_type_ids=tt, attention_mask=am)
    keep = am.bool()
    assert torch.allclose(start[keep], want.start_logits[keep], atol=1e-5) and torch.allclose(end[keep], want.end_logits[keep], atol=1e-5)
    # and back again through this module's own reader, in its own tensor names too
    again, report = models.bert_qa_from_pretrained(str(tmp_path / "out"))
    assert not report["missing"] and all(torch.equal(a, b) for a, b in zip(again.state_dict().values(), model.state_dict().values()))
    models.save_pretrained(model, str(tmp_path / "native"), hf_names=False)
    native, report = models.bert_qa_from_pretrained(str(tmp_path / "native"))
    assert not report["missing"] and torch.equal(native.layers[1].qkv.weight, model.layers[1].qkv.weight)


def test_pretraining_checkpoint_without_a_qa_head_reports_it(tmp_path):
    hm = transformers.BertModel(_hf_model().config)
    hm.save_pretrained(str(tmp_path / "base"))
    _, report = models.bert_qa_from_pretrained(str(tmp_path / "base"))
    assert sorted(report["missing"]) == ["qa_outputs.bias", "qa_outputs.weight"]
    assert os.path.isfile(tmp_path / "base" / "config.json")
