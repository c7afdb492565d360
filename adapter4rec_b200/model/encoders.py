"""Item and user encoders, mirroring Downstream/Text/model/encoders.py."""
import torch
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_

from .layers import BF16, Embedding, Linear, to_2d_bf16
from .modules import TransformerEncoder


class User_Encoder(nn.Module):
    """encoders.py:8-29.  forward(input_embs [B,S,D], log_mask [B,S], local_rank) -> [B,S,D] (bf16)."""

    def __init__(self, item_num, max_seq_len, item_dim, num_attention_heads, dropout, n_layers):
        super().__init__()
        self.transformer_encoder = TransformerEncoder(n_vocab=item_num, n_position=max_seq_len, d_model=item_dim,
                                                      n_heads=num_attention_heads, dropout=dropout, n_layers=n_layers)
        self.apply(self._init_weights)

    def _init_weights(self, module):
        if isinstance(module, (Embedding, Linear)):
            xavier_normal_(module.weight.data)
            if getattr(module, "bias", None) is not None:
                constant_(module.bias.data, 0)

    def forward(self, input_embs, log_mask, local_rank=None):
        # The reference materialises a [B,1,S,S] additive mask (0 / -1e9) from tril(log_mask != 0); the attention
        # kernel rebuilds exactly that mask from log_mask, so the float mask itself is what travels.
        key_mask = log_mask.to(device=input_embs.device, dtype=torch.float32).contiguous()
        return self.transformer_encoder(input_embs, log_mask, key_mask)


class Text_Encoder(nn.Module):
    """encoders.py:38-57: BERT body -> fc(hidden[:, 0]) -> GELU."""

    def __init__(self, bert_model, item_embedding_dim, word_embedding_dim):
        super().__init__()
        self.bert_model = bert_model
        self.fc = Linear(word_embedding_dim, item_embedding_dim)
        self.activate = nn.GELU()

    def forward(self, text):
        batch_size, num_words = text.shape
        num_words = num_words // 2
        text_ids = torch.narrow(text, 1, 0, num_words)
        text_attmask = torch.narrow(text, 1, num_words, num_words)
        if getattr(self.bert_model, "supports_cls_only", False):
            hidden_states = self.bert_model(input_ids=text_ids, attention_mask=text_attmask, cls_only=True)[0]
        else:
            hidden_states = self.bert_model(input_ids=text_ids, attention_mask=text_attmask)[0]
        # CLS rows are read in place by the GEMM's TMA descriptor (row stride = L*H); GELU is its epilogue
        return self.fc(hidden_states[:, 0], act="gelu")


class Bert_Encoder(nn.Module):
    """Item text encoder of the reference (encoders.py:60-99).  An item row is the concatenation of one [ids | mask] block
    per selected text field, in the fixed field order title, abstract, body; every selected field is encoded by the ONE
    Text_Encoder registered as `text_encoders.title` (the reference never builds the other two — SURVEY.md Appendix B-9 —
    and its state_dict keys depend on that) and several fields are averaged.  The hot path has exactly one field."""

    FIELDS = ("title", "abstract", "body")

    def __init__(self, args, bert_model):
        super().__init__()
        if not args.news_attributes:
            raise AssertionError("news_attributes must name at least one text field")
        self.args = args
        widths = {"title": args.num_words_title, "abstract": args.num_words_abstract, "body": args.num_words_body}
        # column span of each field inside an item row; an unselected field occupies no columns
        self.attributes2length, self.attributes2start, column = {}, {}, 0
        for field in self.FIELDS:
            self.attributes2start[field] = column
            self.attributes2length[field] = 2 * widths[field] if field in args.news_attributes else 0
            column += self.attributes2length[field]
        self.newsname = [field for field in self.FIELDS if field in args.news_attributes]
        self.text_encoders = nn.ModuleDict({"title": Text_Encoder(bert_model, args.embedding_dim, args.word_embedding_dim)})

    def forward(self, news):
        encoder = self.text_encoders["title"]
        if len(self.newsname) == 1:          # the hot path: the row IS the field, no slicing copy
            field = self.newsname[0]
            lo, n = self.attributes2start[field], self.attributes2length[field]
            return encoder(news if (lo == 0 and n == news.shape[1]) else news[:, lo:lo + n])
        vectors = [encoder(news[:, self.attributes2start[f]:self.attributes2start[f] + self.attributes2length[f]])
                   for f in self.newsname]
        return torch.stack(vectors, dim=1).float().mean(dim=1).to(vectors[0].dtype)
