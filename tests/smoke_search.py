"""__graft_entry__.smoke(): one small search + value-network invocation on cuda:0, checked against the CPU oracle.
Lives under tests/ because it imports oracle/ (the product package never does)."""
import numpy as np
import torch


def run():
    from oracle import net_oracle as no
    from oracle import oracle as orc
    from alphasnake_zero_b200.engine import Engine
    from alphasnake_zero_b200.net import NativeNet
    G, S, D, B, seed = 8, 4, 8, 16, 11
    eng = Engine(side=11, snakes=S, games=G, seed=seed, max_depth=D, max_breadth=B, softmax_base=2.0, training=True,
                 table_log2=18)
    eng.reset()
    info = eng.search_info()
    tree = torch.full((info["epochs"], info["max_steps"], G * info["P"], S), 255, dtype=torch.uint8, device="cuda")
    q, mv = eng.search(value_fn=None, trace=tree, trace_mode=2)
    games = []
    for gi in range(G):
        g = orc.OracleGame(11, 11, S, 1); g.init_native(seed, gi, 0); g.set_ids(gi, 0); games.append(g)
    agent = orc.OracleAgent(base=2.0, training=True, max_depth=D, max_breadth=B)
    mvh = mv.cpu().numpy()
    root = np.array([mvh[g, s] for g in range(G) for s in range(S)], np.uint8)
    omv, oq = agent.make_moves(games, G, root_turn=0, tree_moves=np.ascontiguousarray(tree.cpu().numpy()),
                               root_moves=root.copy(), replay=True)
    tab, otab = eng.table(), agent.table()
    oo = np.lexsort((otab["keys"][:, 1], otab["keys"][:, 0]))
    assert np.array_equal(tab["keys"], otab["keys"][oo]) and np.array_equal(tab["N"], otab["N"][oo]), "search tables differ"
    assert np.abs(q.cpu().numpy()[:, :S].reshape(-1, 3) - oq).max() < 1e-5
    # value network: tcgen05 path vs the float64 restatement
    w = no.init_weights(11, seed=2, randomize_bn=True)
    X = np.array(games[0].get_states() + games[1].get_states(), np.float32)
    got = NativeNet(w, "cuda", chunk_images=16).forward(torch.from_numpy(X).cuda()).cpu().numpy()
    assert np.abs(got - no.forward(w, X)).max() < 2e-2
    eng.close()
    print("smoke search ok: %d table entries match the oracle; value net within 2e-2" % len(tab["keys"]))
