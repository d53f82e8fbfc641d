"""Continuous batching (fish_speech_b200/scheduler.py) on the GPU.

The contract: a request that shares decode frames with arbitrary neighbours, in any slot, admitted at any
time, returns EXACTLY the tokens the reference-shaped `generate` returns for it alone (integer results:
bit-exact). Greedy requests are additionally pinned to the oracle.
"""
import pytest
import torch

from oracle import lm_oracle as O
from tests.lm_util import build_model, make_prompt

pytestmark = pytest.mark.gpu


def _alone(model, prompt, n, **kw):
    from fish_speech_b200.models.text2semantic.inference import generate

    return generate(model=model, prompt=prompt.cuda(), max_new_tokens=n, **kw).cpu()


def _im_end_weights(cfg, seed, prompt, at_frame, total):
    """Weights for which a greedy run of `prompt` emits <|im_end|> at generated frame `at_frame`."""
    w = O.make_weights(cfg, seed=seed, head_gain=8.0)
    free = O.generate(O.setup(cfg, w), prompt, total, temperature=0.7, top_p=0.7, top_k=1, noise=False)
    T = prompt.size(1)
    w["embeddings.weight"][cfg.im_end_id] = (w["embeddings.weight"][int(free[0, T + at_frame])].float() * 1.5).bfloat16()
    return w


def test_scheduled_requests_equal_generate_alone_greedy_and_sampled():
    from fish_speech_b200.scheduler import ContinuousBatcher, SlotRequest

    cfg = O.tiny_config()
    p0 = make_prompt(cfg, 61, 10)
    w = _im_end_weights(cfg, 61, p0, at_frame=3, total=6)
    st = O.setup(cfg, w)
    # request 0 stops on <|im_end|> (oracle-pinned); the others differ in prompt length, budget and sampling
    specs = [dict(prompt=p0, n=20, temperature=0.7, top_p=0.7, top_k=1, seed=5)]
    for i, (T, n, k) in enumerate([(7, 9, 1), (23, 17, 30), (12, 1, 1), (31, 12, 30), (9, 25, 1), (16, 2, 30),
                                   (5, 14, 30), (40, 8, 1)]):
        specs.append(dict(prompt=make_prompt(cfg, 100 + i, T), n=n, temperature=0.7 if k == 1 else 0.9,
                          top_p=0.7 if k == 1 else 0.85, top_k=k, seed=1000 + 17 * i))

    model = build_model(cfg, w, max_batch=3, debug=False)
    alone = [_alone(model, s["prompt"], s["n"], temperature=s["temperature"], top_p=s["top_p"], top_k=s["top_k"],
                    seed=s["seed"]) for s in specs]
    ref0 = O.generate(st, p0, 20, temperature=0.7, top_p=0.7, top_k=1, noise=False)
    assert torch.equal(alone[0].to(torch.int32), ref0.to(torch.int32)) and alone[0][0, -1].item() == cfg.im_end_id
    for s, a in zip(specs[1:], alone[1:]):  # no early stop in the others unless the model says so
        assert a.shape[1] <= s["prompt"].size(1) + s["n"]

    order = []
    b = ContinuousBatcher(model, max_slots=3, frames_per_poll=4)
    reqs = []
    for i, s in enumerate(specs[:5]):
        reqs.append(b.submit(SlotRequest(prompt=s["prompt"].cuda(), max_new_tokens=s["n"], temperature=s["temperature"],
                                         top_p=s["top_p"], top_k=s["top_k"], seed=s["seed"], tag=i,
                                         on_done=lambda r: order.append(r.tag))))
    b.step()
    b.step()
    for i, s in enumerate(specs[5:], start=5):  # late arrivals join running neighbours
        reqs.append(b.submit(SlotRequest(prompt=s["prompt"].cuda(), max_new_tokens=s["n"], temperature=s["temperature"],
                                         top_p=s["top_p"], top_k=s["top_k"], seed=s["seed"], tag=i,
                                         on_done=lambda r: order.append(r.tag))))
    b.run()
    b.close()
    assert sorted(order) == list(range(len(specs))) and b.idle()
    for i, (r, a) in enumerate(zip(reqs, alone)):
        assert r.done.is_set() and r.result is not None
        got = r.result.cpu()
        assert got.shape == a.shape, (i, got.shape, a.shape)
        assert torch.equal(got, a), (i, (got != a).nonzero()[:4].tolist())
    # slots were reused: 9 requests went through 3 slots, and sharing frames beat running them back to back
    # (same polling granularity: a lone request also decodes in groups of 4 frames)
    sequential = sum(-(-(a.shape[1] - s["prompt"].size(1) - 1) // 4) * 4 for s, a in zip(specs, alone))
    assert b.frames_run < sequential, (b.frames_run, sequential)

    # the engine is handed back in the one-request mode: generate() still gives the same answer
    again = _alone(model, specs[2]["prompt"], specs[2]["n"], temperature=specs[2]["temperature"],
                   top_p=specs[2]["top_p"], top_k=specs[2]["top_k"], seed=specs[2]["seed"])
    assert torch.equal(again, alone[2])


def test_slot_position_in_the_batch_does_not_change_a_sampled_request():
    """The same sampled request in slot 0 alone and in the last slot next to 31 others."""
    from fish_speech_b200.scheduler import ContinuousBatcher, SlotRequest

    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed=71, head_gain=2.5)
    model = build_model(cfg, w, max_batch=32, debug=False)
    probe = dict(prompt=make_prompt(cfg, 7, 11).cuda(), max_new_tokens=12, temperature=0.8, top_p=0.9, top_k=30, seed=99)
    b = ContinuousBatcher(model, max_slots=32, frames_per_poll=6)
    r_alone = b.submit(SlotRequest(**probe))
    b.run()
    others = [b.submit(SlotRequest(prompt=make_prompt(cfg, 200 + i, 6 + i % 9).cuda(), max_new_tokens=5 + i % 13,
                                   temperature=0.8, top_p=0.9, top_k=30, seed=i)) for i in range(31)]
    r_last = b.submit(SlotRequest(**probe))
    b.run()
    b.close()
    assert r_last.slot == 31 and r_alone.slot == 0
    assert all(o.done.is_set() for o in others)
    assert torch.equal(r_alone.result, r_last.result)
    assert len({tuple(o.result[0, -3:].tolist()) for o in others}) > 1  # neighbours really sampled different things


def _grown_prompt(cfg, first: torch.Tensor, generated: torch.Tensor, extra_seed: int, extra: int) -> torch.Tensor:
    """What generate_long does between chunks (inference.py:611-721): the next prompt = the previous prompt, the frames
    just generated (as semantic rows: token id + codes), then new text rows."""
    g = torch.Generator().manual_seed(extra_seed)
    tail = torch.zeros(cfg.num_codebooks + 1, extra, dtype=first.dtype)
    tail[0] = torch.randint(0, cfg.im_end_id, (extra,), generator=g)
    return torch.cat([first, generated.to(first.dtype), tail], dim=1)


def test_chunked_generation_with_prefix_reuse_equals_full_reprefill():
    """every chunk of a long generation extends the previous prompt. With reuse_prefix the engine
    prefills only the new rows (K/V of the old ones is still in the slot); the tokens must equal a full re-prefill."""
    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed=91, head_gain=8.0)
    kw = dict(temperature=0.7, top_p=0.7, top_k=1)
    p1 = make_prompt(cfg, 91, 40)
    reuse_model, fresh_model = build_model(cfg, w, debug=False), build_model(cfg, w, debug=False)
    chunks_reuse, chunks_fresh, prompt = [], [], p1
    for c in range(3):
        a = _alone(reuse_model, prompt, 7, reuse_prefix=True, **kw)
        b = _alone(fresh_model, prompt, 7, **kw)
        chunks_reuse.append(a)
        chunks_fresh.append(b)
        assert torch.equal(a, b), f"chunk {c}: reuse != full re-prefill at {(a != b).nonzero()[:4].tolist()}"
        prompt = _grown_prompt(cfg, prompt, a[:, prompt.size(1):], 300 + c, 9)
    eng = reuse_model.engine
    # chunk 0: nothing to reuse; chunks 1, 2: the whole previous prompt (40, then 40 + 7 + 9 rows)
    assert eng.rows_reused == 40 + 56 and eng.rows_prefilled == 40 + (56 - 40) + (72 - 56)
    # and both follow the oracle on the first chunk (identical except after a bf16 near-tie of the oracle's own logits)
    from tests.lm_util import assert_tokens_match

    traces = []
    ref = O.generate(O.setup(cfg, w), p1, 7, noise=False, traces=traces, **kw)
    assert assert_tokens_match(chunks_fresh[0], ref, traces, cfg, 40, "chunk 0 vs oracle") >= 2


def test_scheduler_shares_a_prompt_prefix_across_slots():
    """Two requests with the same 48-row system/reference prefix: the second one, admitted into another slot while the
    first is still decoding, copies the prefix K/V (fsb_lm_copy_kv) and prefills only its own tail; both return
    exactly what `generate` returns for them alone. A third request reuses the K/V of a RETIRED slot."""
    from fish_speech_b200.scheduler import ContinuousBatcher, SlotRequest

    cfg = O.tiny_config()
    w = O.make_weights(cfg, seed=92, head_gain=8.0)
    model = build_model(cfg, w, max_batch=3, debug=False)
    shared = make_prompt(cfg, 92, 48)
    prompts = [_grown_prompt(cfg, shared, shared[:, :0], 400 + i, 6 + 3 * i) for i in range(3)]
    kw = dict(temperature=0.7, top_p=0.7, top_k=1)
    alone = [_alone(model, p, 10 + 2 * i, **kw) for i, p in enumerate(prompts)]
    eng = model.engine
    b = ContinuousBatcher(model, max_slots=3, frames_per_poll=4)
    eng.rows_reused = eng.rows_prefilled = 0
    r0 = b.submit(SlotRequest(prompt=prompts[0].cuda(), max_new_tokens=10, reuse_prefix=True, **kw))
    b.step()
    r1 = b.submit(SlotRequest(prompt=prompts[1].cuda(), max_new_tokens=12, reuse_prefix=True, **kw))
    b.run()
    assert r0.slot != r1.slot and eng.rows_reused == 48
    r2 = b.submit(SlotRequest(prompt=prompts[2].cuda(), max_new_tokens=14, reuse_prefix=True, **kw))
    b.run()
    b.close()
    assert eng.rows_reused == 96 and eng.rows_prefilled == (48 + 6) + 9 + 12
    for r, a in zip((r0, r1, r2), alone):
        assert torch.equal(r.result.cpu(), a), (r.result.shape, a.shape)
