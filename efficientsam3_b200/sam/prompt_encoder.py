"""PromptEncoder, B200-native.  Module tree / keys of sam3/sam3/sam/prompt_encoder.py (PromptEncoder :12-197,
PositionEmbeddingRandom :200-243).  Points, box corners (labels 2 / 3) and the padding point run on es3_point_embed; a mask prompt
runs on the fused es3_mask_downscale_tokens kernel; without a mask the dense embedding is the broadcast no_mask_embed."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .common import LayerNorm2d


class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=64, scale=None):
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    @torch.no_grad()
    def forward(self, size):
        """[C, h, w] fp32 (prompt_encoder.py:222-233)."""
        h, w = size
        tok = ops.dense_pe(self.positional_encoding_gaussian_matrix.float(), h, w)     # [h*w, C]
        return tok.view(h, w, -1).permute(2, 0, 1)

    def tokens(self, size):
        return ops.dense_pe(self.positional_encoding_gaussian_matrix.float(), size[0], size[1])


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans, activation=nn.GELU):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(4)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(      # parameters kept for checkpoint compatibility
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans), activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self):
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def _tables(self):
        table = torch.cat([e.weight for e in self.point_embeddings], dim=0).detach().float().contiguous()
        return (self.pe_layer.positional_encoding_gaussian_matrix.float(),
                self.not_a_point_embed.weight.detach().float().reshape(-1), table)

    def mask_weights(self):
        """mask_downscaling parameters in the order es3_mask_downscale_tokens takes them."""
        c0, n0, _, c1, n1, _, c2 = self.mask_downscaling
        f = lambda t: t.detach().float().contiguous()
        return (f(c0.weight), f(c0.bias), f(n0.weight), f(n0.bias), f(c1.weight), f(c1.bias), f(n1.weight), f(n1.bias),
                f(c2.weight.reshape(c2.out_channels, -1)), f(c2.bias)), n0.eps

    @torch.no_grad()
    def embed_sparse(self, points, boxes):
        """Sparse prompt tokens (prompt_encoder.py:71-131, 170-181): points (+ padding point when no box) then box corners."""
        gauss, nap, table = self._tables()
        W_, H_ = self.input_image_size[1], self.input_image_size[0]
        parts = []
        if points is not None:
            coords, labels = points
            parts.append(ops.point_embed(coords.float(), labels, gauss, nap, table, W_, H_, pad=(boxes is None)))
        if boxes is not None:
            corners = boxes.float().reshape(-1, 2, 2)
            lab = torch.tensor([[2, 3]], dtype=torch.int32, device=corners.device).expand(corners.shape[0], 2)
            parts.append(ops.point_embed(corners, lab, gauss, nap, table, W_, H_, pad=False))
        if not parts:
            return None
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)

    @torch.no_grad()
    def forward(self, points, boxes, masks):
        """Reference signature (prompt_encoder.py:152-197) -> (sparse [B,N,C] fp32, dense [B,C,h,w] fp32)."""
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        sparse = self.embed_sparse(points, boxes)
        if sparse is None:
            sparse = torch.empty((bs, 0, self.embed_dim), device=self.no_mask_embed.weight.device, dtype=torch.float32)
        h, w = self.image_embedding_size
        if masks is not None:
            wts, eps = self.mask_weights()
            _, tok = ops.mask_downscale_tokens(masks.float(), wts, None, eps, out_bf16=False)
            dense = tok.view(masks.shape[0], h, w, -1).permute(0, 3, 1, 2)
        else:
            dense = self.no_mask_embed.weight.detach().reshape(1, -1, 1, 1).expand(bs, -1, h, w)
        return sparse, dense
