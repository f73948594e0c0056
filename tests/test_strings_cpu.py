"""CPU tests of the host-side string observers against the reference's golden
playthrough and, where oracle/_ref is built, against the compiled reference on random games. The state
fields come from the CPU oracle here; on the GPU the same formatter is fed from the device state
(tests/test_gpu_spiel.py)."""
import numpy as np

from open_spiel_coup_b200.spiel import (DEFAULT_OBS_TYPE, INFO_STATE_OBS_TYPE, PRIVATE_OBS_TYPE, PUBLIC_OBS_TYPE,
                                        IIGObservationType, make_view, observer_string, state_to_string)


def view_from_oracle(s):
    cards = [[(s.players[p].cards[i].value, s.players[p].cards[i].state) for i in range(s.players[p].num_cards)] for p in range(2)]
    hist = [(s.history_player[i], s.history_action[i], s.history_deal_player[i]) for i in range(s.history_len)]
    return make_view(s.turn_number, s.cur_player_move, cards, [s.players[0].coins, s.players[1].coins],
                     [s.players[0].last_action, s.players[1].last_action], hist)


def test_playthrough_strings(oracle, playthrough):
    s = oracle.new_state()
    n = 0
    for st in playthrough["states"]:
        v = view_from_oracle(s)
        if "InformationStateString(0)" in st:      # fully recorded state
            # coup.txt prints ToString() as "# "-prefixed comment lines, which drops trailing blanks
            assert [ln.rstrip() for ln in state_to_string(v).split("\n")] == [ln.rstrip() for ln in st["ToString"].split("\n")]
            for p in (0, 1):
                assert observer_string(v, p, INFO_STATE_OBS_TYPE) == st[f"InformationStateString({p})"]
                assert observer_string(v, p, DEFAULT_OBS_TYPE) == st[f"ObservationString({p})"]
                assert observer_string(v, p, PRIVATE_OBS_TYPE) == st[f"PrivateObservationString({p})"]
            assert observer_string(v, 0, PUBLIC_OBS_TYPE) == st["PublicObservationString"]
            n += 1
        if st.get("action") is not None:
            oracle.apply(s, st["action"])
    assert n == 9


def test_strings_match_reference_on_random_games(oracle, reference):
    rng = np.random.default_rng(11)
    combos = [IIGObservationType(pub, rec, priv) for pub in (0, 1) for rec in (0, 1) for priv in (0, 1, 2)]
    for ep in range(40):
        h = reference.new_state()
        s = oracle.new_state()
        while True:
            v = view_from_oracle(s)
            assert state_to_string(v) == reference.to_string(h)
            for p in (0, 1):
                assert observer_string(v, p, INFO_STATE_OBS_TYPE) == reference.info_state_string(h, p)
                assert observer_string(v, p, DEFAULT_OBS_TYPE) == reference.observation_string(h, p)
                for t in combos:
                    assert observer_string(v, p, t) == reference.observer_string(h, p, t.public_info, t.perfect_recall, t.private_info)
            if reference.is_terminal(h):
                break
            if reference.is_chance(h):
                oc = reference.chance_outcomes(h)
                a = int(rng.choice([x for x, _ in oc], p=[q for _, q in oc]))
            else:
                la = reference.legal_actions(h)
                a = int(la[rng.integers(len(la))])
            reference.apply(h, a)
            oracle.apply(s, a)
        reference.free(h)


def test_action_to_string_matches_reference(reference):
    from open_spiel_coup_b200.spiel import ACTION_NAMES, CARD_NAMES
    for a in range(18):
        assert reference.action_to_string(0, a) == ACTION_NAMES[a]
    for c in range(5):
        assert reference.action_to_string(-1, c) == "Chance drawn card:" + CARD_NAMES[c]


def test_wire_format_loads_in_the_reference(reference, ref_trajectories):
    """trajectories are exported as the reference's own serialisation (State::Serialize,
    spiel.cc:297-311: one action id per line), so the reference's Game::DeserializeState loads them."""
    actions, offsets, _ = ref_trajectories
    for t in range(0, 400, 7):
        acts = actions[offsets[t]:offsets[t + 1]]
        for upto in (len(acts), len(acts) // 2, 4, 0):
            text = "\n".join(str(int(a)) for a in acts[:upto]) + "\n"   # CoupVectorEnv.serialized_states() / CoupState.serialize()
            h = reference.deserialize_state(text)
            g = reference.state_from_actions(acts[:upto])
            assert reference.serialize(h) == text
            assert reference.to_string(h) == reference.to_string(g)
            reference.free(h)
            reference.free(g)
