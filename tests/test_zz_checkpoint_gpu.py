"""GPU: `PTrainer.save_checkpoint` / `resume_or_load` on the real arenas (`pt/engine/trainer.py:104-111,466-495`,
`pt/checkpoint/detection_checkpoint.py`): a second trainer resumed from the file holds bit-identical student,
teacher, momentum and fp16 operand arenas and continues at the next iteration; the file itself carries the
reference's key names (`modelTeacher.` / `modelStudent.` + detectron2 parameter names) and layouts, checked by
loading it into the CPU oracle's reference-named state dict. (Host logic: tests/test_checkpoint_cpu.py.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trainer_checkpoint_resume(cuda, tmp_path):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.synthetic import synthetic_batch
    H, W, K = 192, 272, 8
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.OUTPUT_DIR = str(tmp_path)

    def loader():
        s = 0
        while True:
            lab = synthetic_batch(2, H, W, K, 10 + s)
            unl = synthetic_batch(2, H, W, K, 500 + s, labelled=False)
            yield lab, [dict(d) for d in lab], unl, [dict(d) for d in unl]
            s += 1

    tr = PTrainer(cfg, loader(), device=cuda, seed=3)
    for _ in range(2):
        tr.run_step()
    path = tr.save_checkpoint()
    assert path.endswith("model_0000001.pth")
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert raw["iteration"] == 1 and set(raw) == {"model", "optimizer", "iteration"}
    # the file speaks the reference's names and layouts: the oracle (which takes reference state dicts) loads it
    om = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=0)
    student_sd = {k[len("modelStudent."):]: v for k, v in raw["model"].items() if k.startswith("modelStudent.")}
    assert set(student_sd) == set(om.ref_state_dict())
    om.load_ref_state_dict(student_sd)
    for k, v in om.ref_state_dict().items():
        assert v.shape == student_sd[k].shape and torch.equal(v.detach(), student_sd[k]), k

    tr2 = PTrainer(cfg, loader(), device=cuda, seed=99)
    assert not torch.equal(tr2.model.arena.data, tr.model.arena.data)
    start = tr2.resume_or_load(resume=True)
    torch.cuda.synchronize()
    assert start == 2 and tr2.iter == 2
    for a, b in ((tr.model.arena, tr2.model.arena), (tr.model_teacher.arena, tr2.model_teacher.arena)):
        assert torch.equal(a.data, b.data)
        assert torch.equal(a.half, b.half)  # fp16 GEMM operands re-packed from the loaded masters
    assert torch.equal(tr.model.arena.momentum, tr2.model.arena.momentum)
    assert float(tr.model.arena.momentum.abs().sum()) > 0
    # the teacher differs from the student after an EMA step: the two prefixes did not get mixed up
    assert not torch.equal(tr2.model_teacher.arena.data, tr2.model.arena.data)
    # and the resumed trainer steps
    losses = tr2.run_step()
    torch.cuda.synchronize()
    assert len(losses) == 8 and all(torch.isfinite(v).all() for v in losses.values())

    # weights-only load (resume=False): momentum and iteration stay at their initial values
    cfg3 = c2f_config()
    cfg3.UNSUPNET.BURN_UP_STEP = 0
    cfg3.OUTPUT_DIR = str(tmp_path / "other")
    cfg3.MODEL.WEIGHTS = path
    tr3 = PTrainer(cfg3, loader(), device=cuda, seed=5)
    assert tr3.resume_or_load(resume=False) == 0
    assert torch.equal(tr3.model.arena.data, tr.model.arena.data)
    assert float(tr3.model.arena.momentum.abs().sum()) == 0.0
