"""BERT encoder with a SQuAD question-answering head (Devlin et al. 2018); ``bert_large_config()`` = 24 layers,
hidden 1024, 16 heads, 335 M parameters — the model of the reference's SQuAD example (examples/squad/main.py)."""
from dataclasses import dataclass

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
