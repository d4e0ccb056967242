"""pit Agent -- mirror of code/utils/pit_agent.py:4-28: moves = argmaxs(nnet.v(states)), no search."""


class Agent:

    def __init__(self, nnet, game_and_snake_cnt=None):
        self.nnet = nnet
        self.game_and_snake_cnt = game_and_snake_cnt

    def make_moves(self, states, ids=None):
        if len(states) == 0:
            return []
        V = self.nnet.v(states)
        return self.argmaxs(V)

    def argmaxs(self, Z):
        out = [-1] * len(Z)
        for i in range(len(Z)):
            if Z[i][0] > Z[i][1]:
                out[i] = 0 if Z[i][0] > Z[i][2] else 2
            else:
                out[i] = 1 if Z[i][1] > Z[i][2] else 2
        return out
