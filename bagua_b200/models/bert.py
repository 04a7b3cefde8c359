"""BERT encoder with a SQuAD question-answering head (Devlin et al. 2018); ``bert_large_config()`` = 24 layers,
hidden 1024, 16 heads, 335 M parameters — the model of the reference's SQuAD example (examples/squad/main.py)."""
import json
import os
from dataclasses import asdict, dataclass, fields

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    hidden_dropout_prob: float = 0.1
    layer_norm_eps: float = 1e-12

    @classmethod
    def from_dict(cls, d: dict) -> "BertConfig":
        """From a HuggingFace ``config.json`` dictionary (unknown keys are ignored)."""
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in known})

    @classmethod
    def from_json_file(cls, path: str) -> "BertConfig":
        with open(path) as f:
            return cls.from_dict(json.load(f))

    def to_dict(self) -> dict:
        return dict(asdict(self), model_type="bert", architectures=["BertForQuestionAnswering"], hidden_act="gelu",
                    attention_probs_dropout_prob=0.0)


def bert_large_config() -> BertConfig:
    return BertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.nh = c.num_attention_heads
        self.qkv = nn.Linear(c.hidden_size, 3 * c.hidden_size)
        self.attn_out = nn.Linear(c.hidden_size, c.hidden_size)
        self.ln1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.fc1 = nn.Linear(c.hidden_size, c.intermediate_size)
        self.fc2 = nn.Linear(c.intermediate_size, c.hidden_size)
        self.ln2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.drop = nn.Dropout(c.hidden_dropout_prob)

    def forward(self, x, mask=None):
        B, S, H = x.shape
        q, k, v = self.qkv(x).view(B, S, 3, self.nh, H // self.nh).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
        a = a.transpose(1, 2).reshape(B, S, H)
        x = self.ln1(x + self.drop(self.attn_out(a)))
        return self.ln2(x + self.drop(self.fc2(F.gelu(self.fc1(x)))))


class BertForQuestionAnswering(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.config = c
        self.word = nn.Embedding(c.vocab_size, c.hidden_size)
        self.pos = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.typ = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.ln = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.drop = nn.Dropout(c.hidden_dropout_prob)
        self.layers = nn.ModuleList([BertLayer(c) for _ in range(c.num_hidden_layers)])
        self.qa_outputs = nn.Linear(c.hidden_size, 2)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
        if isinstance(m, nn.Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, start_positions=None, end_positions=None):
        B, S = input_ids.shape
        pos = torch.arange(S, device=input_ids.device).unsqueeze(0)
        tt = token_type_ids if token_type_ids is not None else torch.zeros_like(input_ids)
        x = self.drop(self.ln(self.word(input_ids) + self.pos(pos) + self.typ(tt)))
        mask = None
        if attention_mask is not None:
            mask = attention_mask[:, None, None, :].to(torch.bool)
        for layer in self.layers:
            x = layer(x, mask)
        start_logits, end_logits = self.qa_outputs(x).float().unbind(dim=-1)
        if start_positions is not None:
            loss = (F.cross_entropy(start_logits, start_positions) + F.cross_entropy(end_logits, end_positions)) / 2
            return loss, start_logits, end_logits
        return start_logits, end_logits


# ---------------------------------------------------------------------------------------------------------------------
# HuggingFace checkpoint layout <-> this module (the reference's SQuAD example starts from a `transformers` checkpoint such as
# bert-large-uncased-whole-word-masking, examples/squad/README.md:9-24; there is no network here, so the checkpoint must be a
# local directory: config.json + model.safetensors / pytorch_model.bin)
# ---------------------------------------------------------------------------------------------------------------------
_HF_EMBED = {"word.weight": "embeddings.word_embeddings.weight", "pos.weight": "embeddings.position_embeddings.weight",
             "typ.weight": "embeddings.token_type_embeddings.weight", "ln.weight": "embeddings.LayerNorm.weight", "ln.bias": "embeddings.LayerNorm.bias"}
_HF_LAYER = {"attn_out": "attention.output.dense", "ln1": "attention.output.LayerNorm", "fc1": "intermediate.dense", "fc2": "output.dense", "ln2": "output.LayerNorm"}


def convert_hf_state_dict(hf: dict, num_layers: int) -> dict:
    """HuggingFace ``BertForQuestionAnswering`` / ``BertModel`` / ``BertForPreTraining`` tensors → this module's names. The three
    attention projections are stacked into the single ``qkv`` GEMM (rows: all of Q, then K, then V — the order ``BertLayer.forward``
    splits them in); pooler and pre-training heads are dropped; a missing ``qa_outputs`` (a checkpoint that was never fine-tuned)
    is left to the caller's random initialisation. Old TF-style ``gamma`` / ``beta`` LayerNorm names are accepted."""
    hf = {k.replace(".gamma", ".weight").replace(".beta", ".bias"): v for k, v in hf.items()}
    pre = "bert." if any(k.startswith("bert.") for k in hf) else ""
    out = {}
    for ours, theirs in _HF_EMBED.items():
        out[ours] = hf[pre + theirs]
    for i in range(num_layers):
        base = f"{pre}encoder.layer.{i}."
        for part in ("weight", "bias"):
            out[f"layers.{i}.qkv.{part}"] = torch.cat([hf[f"{base}attention.self.{n}.{part}"] for n in ("query", "key", "value")], dim=0)
            for ours, theirs in _HF_LAYER.items():
                out[f"layers.{i}.{ours}.{part}"] = hf[f"{base}{theirs}.{part}"]
    for part in ("weight", "bias"):
        if f"qa_outputs.{part}" in hf:
            out[f"qa_outputs.{part}"] = hf[f"qa_outputs.{part}"]
    return out


def to_hf_state_dict(sd: dict, num_layers: int) -> dict:
    """Inverse of :func:`convert_hf_state_dict` (``BertForQuestionAnswering`` names), so a fine-tuned model can be handed back to
    ``transformers`` tooling."""
    out = {f"bert.{theirs}": sd[ours] for ours, theirs in _HF_EMBED.items()}
    for i in range(num_layers):
        base = f"bert.encoder.layer.{i}."
        for part in ("weight", "bias"):
            q, k, v = sd[f"layers.{i}.qkv.{part}"].chunk(3, dim=0)
            out[f"{base}attention.self.query.{part}"], out[f"{base}attention.self.key.{part}"], out[f"{base}attention.self.value.{part}"] = q, k, v
            for ours, theirs in _HF_LAYER.items():
                out[f"{base}{theirs}.{part}"] = sd[f"layers.{i}.{ours}.{part}"]
    out["qa_outputs.weight"], out["qa_outputs.bias"] = sd["qa_outputs.weight"], sd["qa_outputs.bias"]
    return {k: v.detach().clone().contiguous() for k, v in out.items()}


def _read_weights(directory: str) -> dict:
    st, pt = os.path.join(directory, "model.safetensors"), os.path.join(directory, "pytorch_model.bin")
    if os.path.isfile(st):
        from safetensors.torch import load_file

        return load_file(st)
    if os.path.isfile(pt):
        return torch.load(pt, map_location="cpu")
    raise FileNotFoundError(f"neither model.safetensors nor pytorch_model.bin in {directory}")


def bert_qa_from_pretrained(directory: str, config: "BertConfig" = None):
    """``BertForQuestionAnswering`` from a local HuggingFace-format directory. Returns ``(model, report)`` where report lists what the
    checkpoint did not provide (typically the QA head of a pre-trained-only checkpoint)."""
    cfg = config or BertConfig.from_json_file(os.path.join(directory, "config.json"))
    model = BertForQuestionAnswering(cfg)
    raw = _read_weights(directory)
    ours = convert_hf_state_dict(raw, cfg.num_hidden_layers) if any("encoder.layer." in k for k in raw) else raw   # already this module's names
    res = model.load_state_dict(ours, strict=False)
    return model, {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys)}


def save_pretrained(model: "BertForQuestionAnswering", directory: str, hf_names: bool = True) -> None:
    """config.json + model.safetensors in ``directory`` (HuggingFace tensor names by default)."""
    from safetensors.torch import save_file

    os.makedirs(directory, exist_ok=True)
    cfg = model.config
    with open(os.path.join(directory, "config.json"), "w") as f:
        json.dump(cfg.to_dict(), f, indent=1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    save_file(to_hf_state_dict(sd, cfg.num_hidden_layers) if hf_names else {k: v.contiguous() for k, v in sd.items()}, os.path.join(directory, "model.safetensors"))
