"""Diagnostic (not a test): per-tensor gradient errors of the training path vs oracle autograd."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tests.test_gpu_train as T

mode = sys.argv[1] if len(sys.argv) > 1 else "xe"
drop = len(sys.argv) > 2 and sys.argv[2] == "1"
m, sd = T._model()
fc, att, cpts, sentis, labels, caps, lengths = T._inputs()
masks = T._masks(mode, T.T1 - 1) if drop else None
m.train(drop)
m.dropout_override = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in masks.items()} if masks else None
m.zero_grad()
if mode == "xe":
    pred = m(fc.cuda(), att.cuda(), cpts.cuda(), caps.cuda(), labels.cuda(), 0.0, mode="xe")
    loss = T._xe_loss(pred, caps[:, 1:].cuda(), lengths) + torch.nn.functional.mse_loss(m.cpt_feats, m.fc_feats.detach())
    def ref_loss(p):
        f = T.O.prologue(p, fc, att, cpts, None, labels, masks=masks)
        lp = T.O.teacher_forced(p, f, caps, masks=masks)
        return T._xe_loss(lp, caps[:, 1:], lengths) + torch.nn.functional.mse_loss(f["cpt_feats"], f["fc_embedded"].detach())
elif mode == "rl":
    TT = 6
    masks = T._masks("rl", TT) if drop else None
    m.dropout_override = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in masks.items()} if masks else None
    g = torch.Generator().manual_seed(9)
    noise = -torch.log(-torch.log(torch.rand(TT, T.B, T.V, generator=g).clamp_min(1e-9)))
    rewards = torch.randn(T.B, TT, generator=g)
    seq, lps, smask = m(fc.cuda(), att.cuda(), cpts.cuda(), sentis.cuda(), labels.cuda(), TT, 0, mode="rl", noise=noise)
    loss = -(lps * smask * rewards.cuda()).sum() / smask.sum() + torch.nn.functional.mse_loss(m.cpt_feats, m.fc_feats.detach())
    seq_c, smask_c = seq.cpu(), smask.cpu()
    print("seq", seq_c.tolist())
    def ref_loss(p):
        f = T.O.prologue(p, fc, att, cpts, sentis, labels, masks=masks)
        inputs = torch.cat([torch.full((T.B, 1), 1, dtype=torch.long), seq_c], dim=1)
        lp = T.O.teacher_forced(p, f, inputs, masks=masks)
        executed = smask_c.sum(0, keepdim=True).gt(0).float()
        chosen = lp.gather(2, seq_c.unsqueeze(2)).squeeze(2) * executed
        return -(chosen * smask_c * rewards).sum() / smask_c.sum() + torch.nn.functional.mse_loss(f["cpt_feats"], f["fc_embedded"].detach())
else:
    pred = m(caps.cuda(), cpts.cuda(), sentis.cuda(), labels.cuda(), 0.0, mode="seq2seq")
    loss = T._xe_loss(pred, caps[:, 1:].cuda(), lengths)
    def ref_loss(p):
        f = T.O.prologue(p, None, None, cpts, sentis, labels, seq2seq=True, masks=masks)
        return T._xe_loss(T.O.teacher_forced(p, f, caps, masks=masks), caps[:, 1:], lengths)
loss.backward()
torch.cuda.synchronize()
ref, grads = T._oracle_grads(sd, ref_loss)
print("loss", float(loss), ref)
print("senti words", sentis.tolist(), "cpts", cpts.tolist())
for name, prm in m.named_parameters():
    got = prm.grad.detach().cpu() if prm.grad is not None else torch.zeros_like(grads[name])
    r = grads[name]
    err = (got - r).abs()
    print("%-42s ref max %.3e  err max %.3e  rel %.2e  argmax %s" % (name, float(r.abs().max()), float(err.max()),
          float(err.max()) / (float(r.abs().max()) + 1e-30), tuple(int(i) for i in torch.nonzero(err == err.max())[0])))
# ReLU-boundary check: units whose activity differs between the GPU prologue and the oracle
if mode == "rl":
    with torch.no_grad():
        t, _ = m.prologue(fc.cuda(), att.cuda(), cpts.cuda(), sentis.cuda(), labels.cuda(), dropout=m.dropout_override)
        f = T.O.prologue(sd, fc, att, cpts, sentis, labels, masks=masks)
        pre_sw = torch.nn.functional.linear(f["sw"], sd["senti2att.0.weight"], sd["senti2att.0.bias"])
        pre_pa = torch.nn.functional.linear(f["att"], sd["att2att.0.weight"], sd["att2att.0.bias"])
    for name, gpu_active, pre in (("p_sw", t["p_sw"].cpu() < 1.0, pre_sw), ("p_att", t["p_att"].cpu() < 1.0, pre_pa)):
        mism = gpu_active != (pre > 0)
        idx = torch.nonzero(mism)
        print(name, "gate mismatches:", int(mism.sum()), [(tuple(int(x) for x in i), float(pre[tuple(i)])) for i in idx[:8]])
