"""Arena (reference: alpha/Arena.py:4-123): pit two agents on the single-player packing game.

An agent is a function ``canonical board -> action``.  The game has one player, so a game is
played by ONE agent (``playGame`` only ever asks ``player1``, Arena.py:37-49) and "pitting" means:
``num / 2`` games by player 1, then the players swap and ``num / 2`` games by the other one
(Arena.py:84-121).  A game's result is ``getGameEnded(board, 1)``: 1 for a perfect packing, the
filled fraction otherwise — so everything but a perfect packing counts as a draw, exactly as
the reference counts it.
"""
import numpy as np


class Arena(object):
    def __init__(self, player1, player2, game, display=None):
        self.player1, self.player2, self.game, self.display = player1, player2, game, display

    def playGame(self, verbose=False):
        """One episode by ``player1`` on a fresh instance -> getGameEnded(final board, 1)."""
        player = self.player1
        cur = 1
        board, vis_state = self.game.getInitBoard()
        turn = 0
        ended = self.game.getGameEnded(board, cur)
        while ended == 0:
            turn += 1
            if verbose:
                assert self.display
                print("Turn ", str(turn), "Player ", str(cur))
                self.display(board)
            canonical = self.game.getCanonicalForm(board, cur)
            action = player(canonical)
            valids = self.game.getValidMoves(canonical, 1)
            if valids[action] == 0:
                raise AssertionError("player chose the invalid action %d" % action)   # Arena.py:51-53
            board, cur, vis_state = self.game.getNextState(board, cur, action, vis_state)
            ended = self.game.getGameEnded(board, cur)
        if verbose:
            assert self.display
            print("Game over: Turn ", str(turn), "Result ", str(self.game.getGameEnded(board, 1)))
            self.display(board)
        return self.game.getGameEnded(board, 1)

    def playGames(self, num, verbose=False):
        """-> (oneWon, twoWon, draws) over int(num / 2) games by each player (Arena.py:67-123)."""
        half = int(num / 2)
        one = two = draws = 0
        for _ in range(half):
            result = self.playGame(verbose=verbose)
            if result == 1:
                one += 1
            elif result == -1:
                two += 1
            else:
                draws += 1
        self.player1, self.player2 = self.player2, self.player1
        for _ in range(half):
            result = self.playGame(verbose=verbose)
            if result == -1:
                one += 1
            elif result == 1:
                two += 1
            else:
                draws += 1
        return one, two, draws


def mcts_player(mcts, temp=0):
    """The agent the reference's Coach pits (Coach.py:116-117): argmax of getActionProb."""
    return lambda board: int(np.argmax(mcts.getActionProb(board, temp=temp)))
