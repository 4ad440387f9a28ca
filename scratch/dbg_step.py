import numpy as np, torch
from mladversarialobjectdetection_b200 import ops, synth, victim
from mladversarialobjectdetection_b200.attacker import PatchAttacker
from mladversarialobjectdetection_b200.ragged import RaggedBoxes
from oracle import objective, patcher, step as ostep, tfops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
F=np.float32
H, P, B = 128, 32, 3
model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=H, seed=5)
model.class_net.out_pw.bias.data.view(9, 90)[:, 0] += 6.0
cpu_model = victim.get_victim_model("efficientdet-d0", device="cpu", image_size=H, seed=5)
cpu_model.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
bt = synth.make_batch(B, H, H, seed=77, max_boxes=3, min_boxes=1)
att = PatchAttacker(model, patch_size=P, device="cuda", seed=3)
patch0 = att._patch.cpu().numpy().copy()
images = torch.from_numpy(bt.images).cuda()
boxes = RaggedBoxes(torch.from_numpy(bt.boxes).cuda(), torch.from_numpy(bt.offsets).cuda())
tr = (ops.params_to_tensor(bt.params, "cuda"), torch.from_numpy(bt.print_wb).cuda())
att.always_first_pass = False
patched = att._patcher([boxes, images], transforms=tr)
patched.requires_grad_(True)
cls_outputs, M, argmax, ncand, sctx = att.second_pass(patched)
dcls, dscale, data_loss = ops.score_max_backward(sctx, att._scale_regressor)
torch.autograd.backward(cls_outputs, dcls)
Gg = patched.grad
print("grad contiguous", Gg.is_contiguous(), Gg.shape, Gg.stride())
gp = att._patcher.backward(Gg).cpu().numpy()
bx, pr = bt.ragged()
ref = ostep.attack_step(cpu_model, patch0, 0.4, bt.images, bx, pr, bt.print_wb, objective.anchor_boxes(H), first_pass=False)
print("patched equal", np.array_equal(patched.detach().cpu().numpy(), ref["patched"]))
print("M", M.cpu().numpy(), ref["max_scores"], argmax.cpu().numpy())
G = Gg.cpu().numpy(); Gr = ref["grad_images"]
print("G rel", np.linalg.norm(G-Gr)/np.linalg.norm(Gr), np.abs(Gr).max(), np.abs(G).max())
for b in range(B): print(" img", b, np.linalg.norm(G[b]-Gr[b])/np.linalg.norm(Gr[b]))
# oracle backward with GPU's G
_,_,states = patcher.patcher_forward(patch0, bt.images, bx, pr, bt.print_wb, 0.4)
g_or = patcher.patcher_backward(G, patch0, bt.print_wb, states)
print("kernel-vs-oracle on same G:", np.linalg.norm(gp-g_or)/np.linalg.norm(g_or))
gref = ref["grad_patch"] - F(1e-5)*tfops.total_variation(patch0)[1]
print("step rel", np.linalg.norm(gp-gref)/np.linalg.norm(gref))
# victim fwd compare
with torch.no_grad():
    cg,_ = model(patched.detach()); cc,_ = cpu_model(patched.detach().cpu())
for a,b in zip(cg,cc): print(" cls diff", float((a.cpu()-b).abs().max()), float(b.abs().max()))
