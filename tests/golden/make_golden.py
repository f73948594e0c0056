"""Regenerates the golden fixtures in tests/golden/ from the reference. Run in the build container
(needs /root/reference and oracle/_ref/libcoup_ref.so):

    python tests/golden/make_golden.py

Outputs (committed; they travel to the GPU box, /root/reference does not):
  playthrough_coup.json  parsed from the reference's own golden file
                         open_spiel/integration_tests/playthroughs/coup.txt (every state it records:
                         current player, legal actions, chance outcomes, rewards, returns, and the
                         info-state / observation tensors of both players as non-zero lists) plus the
                         string observers for the "next"-tier parity tests.
  ref_trajectories.npz   2 000 trajectories sampled by running the compiled reference (uniform
                         random play, every 20th game steered into Exchange/Pass loops so that the
                         91-move truncation and the post-ExchangeReturn deck quirk are covered), with the
                         reference's per-state record for every prefix (oracle/ref_harness.cc RefTraceRec).
  ref_tensors.npz        full fp32 info-state / observation tensors of both players at 400 sampled
                         states of those trajectories (dense check on top of the hashes).
kat_scenarios.json (the 14 scenario tests of open_spiel/games/coup_test.cc:41-556 as data) is a
hand transcription and is NOT regenerated here; this script only re-validates it against the reference.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.bindings import INFO_SIZE, OBS_SIZE, Reference, build_ref  # noqa: E402

PLAYTHROUGH = "/root/reference/open_spiel/integration_tests/playthroughs/coup.txt"

INFO_LAYOUT = [("player", 2), ("p1_cards", 20), ("p2_cards", 20), ("cur_move_player", 2),
               ("cards_state", 16), ("coins", 2), ("history", 135 * 18)]
OBS_LAYOUT = INFO_LAYOUT[:-1] + [("last_action", 36)]


def _bits(s):
    return [1.0 if ch == "◉" else 0.0 for ch in s if ch in "◉◯"]


def _parse_tensor(pieces, layout):
    flat = []
    for name, size in layout:
        lines = pieces[name]
        if name == "coins":
            vals = [float(x) for x in re.findall(r"[-0-9.]+", lines[0])]
        elif name == "cards_state":
            # 3-D [player][slot][2] is printed as one line per slot with the players side by side
            rows = [ln.split() for ln in lines if ln.strip()]
            vals = [0.0] * 16
            for slot, blocks in enumerate(rows):
                for p, blk in enumerate(blocks):
                    b = _bits(blk)
                    vals[(p * 4 + slot) * 2 + 0] = b[0]
                    vals[(p * 4 + slot) * 2 + 1] = b[1]
        else:
            vals = []
            for ln in lines:
                vals += _bits(ln)
        assert len(vals) == size, (name, len(vals), size)
        flat += vals
    return flat


def parse_playthrough(path):
    text = open(path, encoding="utf-8").read().split("\n")
    header = {}
    states = []
    cur = None
    i = 0
    tensor_re = re.compile(r"^(InformationStateTensor|ObservationTensor)\((\d)\)\.(\w+)(:| =)(.*)$")
    while i < len(text):
        ln = text[i]
        m = re.match(r"^# State (\d+)$", ln)
        if m:
            cur = {"index": int(m.group(1)), "to_string": [], "pieces": {}}
            states.append(cur)
            i += 1
            while i < len(text) and text[i].startswith("#") and not text[i].startswith("# Apply action") \
                    and not re.match(r"^# State \d+$", text[i]):
                cur["to_string"].append(text[i][2:] if len(text[i]) > 1 else "")
                i += 1
            continue
        m = re.match(r"^action: (\d+)$", ln)
        if m and cur is not None:
            cur["action"] = int(m.group(1))
        tm = tensor_re.match(ln)
        if tm and cur is not None:
            kind, player, name, _, rest = tm.groups()
            lines = [rest]
            i += 1
            while i < len(text) and text[i] and (text[i][0] in " ◉◯"):
                lines.append(text[i])
                i += 1
            cur["pieces"].setdefault((kind, int(player)), {})[name] = lines
            continue
        m = re.match(r"^(\w+)\((\d*)\) = (.*)$", ln)
        if m:
            key, arg, val = m.groups()
            tgt = cur if cur is not None else header
            name = key if arg == "" else f"{key}({arg})"
            tgt.setdefault("fields", {})[name] = val if cur is not None else val
        i += 1
    out_states = []
    for st in states:
        f = st.get("fields", {})
        rec = {"index": st["index"], "action": st.get("action")}
        if st["to_string"]:
            rec["ToString"] = "\n".join(st["to_string"]) + "\n"
        for key in ("IsTerminal", "IsChanceNode"):
            if key in f:
                rec[key] = f[key] == "True"
        if "CurrentPlayer" in f:
            rec["CurrentPlayer"] = int(f["CurrentPlayer"])
        for key in ("History", "LegalActions", "Rewards", "Returns"):
            if key in f:
                rec[key] = json.loads(f[key])
        if "ChanceOutcomes" in f:
            rec["ChanceOutcomes"] = [[int(a), float(p)] for a, p in re.findall(r"\((\d+), ([0-9.e-]+)\)", f["ChanceOutcomes"])]
        for key in ("InformationStateString(0)", "InformationStateString(1)", "ObservationString(0)",
                    "ObservationString(1)", "PublicObservationString", "PrivateObservationString(0)",
                    "PrivateObservationString(1)"):
            if key in f:
                rec[key] = json.loads(f[key])
        for (kind, player), pieces in st["pieces"].items():
            layout = INFO_LAYOUT if kind == "InformationStateTensor" else OBS_LAYOUT
            flat = _parse_tensor(pieces, layout)
            rec[f"{kind}({player})"] = [[k, v] for k, v in enumerate(flat) if v != 0.0]
        out_states.append(rec)
    return {"header": header.get("fields", {}), "states": out_states}


def sample_trajectories(ref, n, seed):
    rng = np.random.default_rng(seed)
    trajs = []
    for ep in range(n):
        h = ref.new_state()
        acts = []
        steer = ep % 20 == 0
        while not ref.is_terminal(h):
            if ref.is_chance(h):
                oc = ref.chance_outcomes(h)
                a = int(rng.choice([x for x, _ in oc], p=[p for _, p in oc]))
            else:
                la = ref.legal_actions(h)
                a = int(la[rng.integers(len(la))])
                if steer:  # prefer Exchange / Pass / ExchangeReturn: long games, deck-quirk coverage
                    for x in la:
                        if x in (5, 9, 17, 16, 15):
                            a = x
                            if x in (9, 5):
                                break
            ref.apply(h, a)
            acts.append(a)
        ref.free(h)
        trajs.append(np.array(acts, np.uint8))
    return trajs


def main():
    assert build_ref(), "reference build unavailable"
    ref = Reference()
    pt = parse_playthrough(PLAYTHROUGH)
    # Self-check of the parser against the compiled reference along the playthrough history.
    hist = [s["action"] for s in pt["states"] if s.get("action") is not None]
    h = ref.new_state()
    n_checked = 0
    for st in pt["states"]:
        for p in (0, 1):
            key = f"InformationStateTensor({p})"
            if key in st:
                dense = np.zeros(INFO_SIZE, np.float32)
                for k, v in st[key]:
                    dense[k] = v
                assert np.array_equal(dense, ref.info_state(h, p)), (st["index"], key)
                n_checked += 1
            key = f"ObservationTensor({p})"
            if key in st:
                dense = np.zeros(OBS_SIZE, np.float32)
                for k, v in st[key]:
                    dense[k] = v
                assert np.array_equal(dense, ref.observation(h, p)), (st["index"], key)
                n_checked += 1
        if "LegalActions" in st:
            assert st["LegalActions"] == ref.legal_actions(h)
        if st.get("action") is not None:
            ref.apply(h, st["action"])
    ref.free(h)
    print(f"playthrough: {len(pt['states'])} states, history {hist}, {n_checked} tensors re-checked vs reference")
    with open(os.path.join(HERE, "playthrough_coup.json"), "w", encoding="utf-8") as f:
        json.dump(pt, f, ensure_ascii=False, separators=(",", ":"))

    trajs = sample_trajectories(ref, 2000, seed=20261018)
    flat = np.concatenate(trajs)
    off = np.concatenate([[0], np.cumsum([len(t) for t in trajs])]).astype(np.int64)
    recs, bad = ref.trace_batch(flat, off, 8)
    assert bad == 0
    n_trunc = sum(len(t) > 90 for t in trajs)
    print(f"trajectories: {len(trajs)}, moves {len(flat)}, truncated {n_trunc}, max len {max(map(len, trajs))}")
    np.savez_compressed(os.path.join(HERE, "ref_trajectories.npz"), actions=flat, offsets=off, records=recs)

    rng = np.random.default_rng(7)
    picks = []
    info = np.zeros((400, 2, INFO_SIZE), np.float32)
    obs = np.zeros((400, 2, OBS_SIZE), np.float32)
    for k in range(400):
        t = int(rng.integers(len(trajs)))
        upto = int(rng.integers(len(trajs[t]) + 1))
        picks.append((t, upto))
        h = ref.state_from_actions(trajs[t][:upto])
        for p in (0, 1):
            info[k, p] = ref.info_state(h, p)
            obs[k, p] = ref.observation(h, p)
        ref.free(h)
    np.savez_compressed(os.path.join(HERE, "ref_tensors.npz"), picks=np.array(picks, np.int64),
                        info=info.astype(np.uint8), obs=obs.astype(np.uint8))

    # Re-validate the hand-transcribed KATs against the reference.
    kats = json.load(open(os.path.join(HERE, "kat_scenarios.json")))
    for kat in kats["scenarios"]:
        h = ref.new_state()
        for step in kat["steps"]:
            if "apply" in step:
                assert ref.apply(h, step["apply"]) == 0, (kat["name"], step)
            exp = step.get("expect", {})
            if "current_player" in exp:
                assert ref.current_player(h) == exp["current_player"], (kat["name"], step)
            if "legal_actions" in exp:
                assert ref.legal_actions(h) == exp["legal_actions"], (kat["name"], step)
            if "coins" in exp:
                assert [ref.coins(h, 0), ref.coins(h, 1)] == [c if c is not None else ref.coins(h, i) for i, c in enumerate(exp["coins"])], (kat["name"], step)
            if "is_terminal" in exp:
                assert ref.is_terminal(h) == exp["is_terminal"], (kat["name"], step)
            if "rewards" in exp:
                assert ref.rewards(h) == exp["rewards"], (kat["name"], step)
            if "returns" in exp:
                assert ref.returns(h) == exp["returns"], (kat["name"], step)
        ref.free(h)
    print(f"kat_scenarios.json: {len(kats['scenarios'])} scenarios re-validated against the reference")


if __name__ == "__main__":
    main()
