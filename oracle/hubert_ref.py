"""CPU oracle (TEST INFRASTRUCTURE, not product code): a functional fp32 restatement of the
HuBERT forward that `sylber.Segmenter` runs (reference: sylber/model/sylber.py:41,122).

The arithmetic lives in a third-party dependency of the reference, `transformers.HubertModel`
(pinned 4.45.2 in the reference's requirements.txt:59; 5.5.0 installed in this image).  This
file restates that forward with plain torch CPU ops from a flat state_dict so that
  * parity tests can compare every stage of the CUDA path against it,
  * `bench.py --impl reference` / `cpu_baseline` have something that travels to the GPU box
    (`/root/reference` does not), built from the same torch CPU kernels the reference dispatches to.
It is pinned against `transformers.HubertModel` itself in tests/test_oracle_model.py and against
fixtures produced by the unmodified reference in tests/golden/ (see tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu/reference legs may import this module.

HF line references are to transformers/models/hubert/modeling_hubert.py (5.5.0).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
CONV_DIM = 512
HIDDEN = 768
HEADS = 12
FFN = 3072
POS_K = 128
POS_GROUPS = 16
LN_EPS = 1e-5


def conv_out_lengths(n_samples):
    """Frame count after each of the 7 'valid' convs (HF:675-688)."""
    out = []
    n = n_samples
    for k, s in zip(CONV_KERNEL, CONV_STRIDE):
        n = (n - k) // s + 1
        out.append(n)
    return out


def num_frames(n_samples):
    return conv_out_lengths(n_samples)[-1]


def pos_conv_weight(sd):
    """Fold weight-norm (dim=2: the norm runs over (out, in) for each of the 128 taps, HF:78)."""
    pre = "encoder.pos_conv_embed.conv."
    if pre + "parametrizations.weight.original0" in sd:
        g = sd[pre + "parametrizations.weight.original0"]
        v = sd[pre + "parametrizations.weight.original1"]
    elif pre + "weight_g" in sd:
        g = sd[pre + "weight_g"]
        v = sd[pre + "weight_v"]
    else:
        return sd[pre + "weight"]
    return v * (g / v.norm(2, dim=(0, 1), keepdim=True))


def feature_encoder(sd, wav, stages=None, q=None):
    """7 conv layers (HF:203-213).  wav (B, T_samp) -> (B, 512, L6).  `q` optionally rounds GEMM operands."""
    q = q or (lambda t: t)
    h = wav[:, None, :]
    p = "feature_extractor.conv_layers."
    h = F.conv1d(h, sd[p + "0.conv.weight"], stride=CONV_STRIDE[0])          # HF:172 (fp32 CUDA-core op in our build)
    if stages is not None:
        stages["conv0_raw"] = h
    h = F.group_norm(h, CONV_DIM, sd[p + "0.layer_norm.weight"], sd[p + "0.layer_norm.bias"], eps=1e-5)  # HF:173
    h = F.gelu(h)                                                            # HF:174
    if stages is not None:
        stages["conv0"] = h
    for i in range(1, 7):
        h = F.conv1d(q(h), q(sd[p + f"{i}.conv.weight"]), stride=CONV_STRIDE[i])  # HF:122
        h = F.gelu(h)                                                        # HF:123
        if stages is not None:
            stages[f"conv{i}"] = h
    return h


def frame_mask(n_samples, T):
    """HF:690-700 reduced to what it computes: a prefix mask of valid frames per utterance."""
    valid = torch.tensor([num_frames(int(n)) for n in n_samples])
    return torch.arange(T)[None, :] < valid[:, None]


def encoder_layer(sd, i, h, key_bias, q=None, q_attn=None):
    """Post-LN transformer layer (HF:388-405).  `q` rounds the operands of the linears, `q_attn` (default: q) those of
    Q.K^T and P.V - precision studies only."""
    q = q or (lambda t: t)
    qa = q_attn or q
    p = f"encoder.layers.{i}."
    B, T, _ = h.shape

    def lin(x, name):
        return F.linear(q(x), q(sd[p + name + ".weight"]), sd[p + name + ".bias"])

    qh = lin(h, "attention.q_proj").view(B, T, HEADS, -1).transpose(1, 2)
    kh = lin(h, "attention.k_proj").view(B, T, HEADS, -1).transpose(1, 2)
    vh = lin(h, "attention.v_proj").view(B, T, HEADS, -1).transpose(1, 2)
    s = torch.matmul(qa(qh), qa(kh).transpose(2, 3)) * (qh.shape[-1] ** -0.5)  # HF:248
    if key_bias is not None:
        s = s + key_bias
    pr = torch.softmax(s, dim=-1)
    a = torch.matmul(qa(pr), qa(vh)).transpose(1, 2).reshape(B, T, HIDDEN)
    a = lin(a, "attention.out_proj")
    h = F.layer_norm(h + a, (HIDDEN,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], LN_EPS)
    f = F.gelu(lin(h, "feed_forward.intermediate_dense"))
    f = lin(f, "feed_forward.output_dense")
    h = F.layer_norm(h + f, (HIDDEN,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], LN_EPS)
    return h


@torch.no_grad()
def hubert_forward(sd, wav, n_samples=None, n_layers=9, stages=None, q=None, q_conv=None, q_attn=None):
    """last_hidden_state of HubertModel.forward (HF:889-958) in eval mode.

    sd        flat fp32 state_dict with HubertModel keys
    wav       (B, T_samp) fp32, zero padded to the batch max
    n_samples per-utterance valid sample counts, or None for "no attention mask"
    stages    optional dict that receives per-stage activations
    q, q_conv optional operand-rounding hooks (encoder GEMMs / conv1-6) used only for precision studies
    """
    qq = q or (lambda t: t)
    sd = {k: v.float() for k, v in sd.items()}
    feats = feature_encoder(sd, wav, stages, q_conv if q_conv is not None else q).transpose(1, 2)              # HF:931-932  (B,T,512)
    B, T, _ = feats.shape
    h = F.layer_norm(feats, (CONV_DIM,), sd["feature_projection.layer_norm.weight"],
                     sd["feature_projection.layer_norm.bias"], LN_EPS)       # HF:228
    h = F.linear(qq(h), qq(sd["feature_projection.projection.weight"]), sd["feature_projection.projection.bias"])
    if stages is not None:
        stages["proj"] = h
    key_bias = None
    if n_samples is not None:
        m = frame_mask(n_samples, T)
        h = h * m[:, :, None]                                                # HF:429-432
        if not bool(m.all()):
            key_bias = torch.zeros(B, 1, 1, T).masked_fill(~m[:, None, None, :], float("-inf"))  # HF:434-438
    w = pos_conv_weight(sd)
    pos = F.conv1d(qq(h.transpose(1, 2)), qq(w), sd["encoder.pos_conv_embed.conv.bias"],
                   padding=POS_K // 2, groups=POS_GROUPS)[:, :, :-1]         # HF:87-88, trim HF:98-103
    pos = F.gelu(pos).transpose(1, 2)
    if stages is not None:
        stages["pos"] = pos
    h = F.layer_norm(h + pos, (HIDDEN,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], LN_EPS)
    if stages is not None:
        stages["enc_in"] = h
    for i in range(n_layers):
        h = encoder_layer(sd, i, h, key_bias, q, q_attn)
        if stages is not None:
            stages[f"layer{i}"] = h
    return h
