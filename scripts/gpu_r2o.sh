# round 2, call o: how the adjoint's 15 ms split between pass 1 and pass 2 (volpathsimple primal render = pass 1 without the
# records), and the A/B of the vertex-record prefetch
TAG=${1:-r2o}
mkdir -p gpurun_out
python - > gpurun_out/${TAG}_pass1.json <<PY
import json, sys, torch
sys.path.insert(0, '.')
from recursive_control_variates_b200.runtime import Context, tea32
from recursive_control_variates_b200.scene import SceneConfig
from recursive_control_variates_b200.synthetic import make_medium
ctx = Context(0); ctx.set_scene(SceneConfig(width=512, height=512).to_desc())
for p in (0.0, 5e-3):
    s, a = make_medium(256, 'cube', p); ctx.rotate(torch.from_numpy(s).cuda(), torch.from_numpy(a).cuda())
acc = torch.zeros(512, 512, 4, device='cuda'); g = torch.full((512, 512, 3), 1.0 / (512 * 512 * 3), device='cuda')
fw = ctx.film_weights(tea32(17, 1)[0], 64)
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {'vps_primal_ms': t(lambda: ctx.render_adjoint_primal_accum(tea32(17, 1)[0], 64, out=acc)),
       'vps_backward_ms': t(lambda: ctx.render_backward(tea32(17, 1)[0], 64, g, film_w=fw)),
       'twostate_ms': t(lambda: ctx.render_twostates_accum(17, 64)),
       'single_state_primal_ms': t(lambda: ctx.render_primal_accum(1, 17, 64, out=acc))}
print(json.dumps(out))
PY
cat gpurun_out/${TAG}_pass1.json
bash scripts/gpu_ab_variant.sh ${TAG} pfv
