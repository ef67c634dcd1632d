"""Search-tree node of the flat Monte-Carlo search and its exports.

Interface of the reference's ``State`` (engine/state.py:63-153): the attributes callers
touch (``board``, ``tiles``, ``parent``, ``children``, ``score``, ``tile_placed``,
``solution_tiles_order``, ``uuid``) and the renderers the Django view layer consumes
(``render_to_json``, ``render_children``, ``to_json``).  On the device a state is never an
object: it is a packed board plus a tile table; these nodes are built
on the host from the per-depth child scores a search returns (mcts.CustomMCTS._rebuild_tree).
"""
import itertools
from collections import OrderedDict

import numpy as np

ORIENTATIONS = 2

_ids = itertools.count(1)


def sort_key(node):
    return node.score


class State(object):

    def __init__(self, board, tiles, parent=None, solution_tiles_order=None):
        # a node owns private copies: callers mutate boards in place (destroy_state=True)
        self.board = np.array(board, copy=True)
        self.tiles = list(tiles)
        self.parent = parent
        self.uuid = next(_ids)
        self.children = []
        self.score = None
        self.tile_placed = None
        inherited = getattr(parent, "solution_tiles_order", None) if parent is not None else None
        self.solution_tiles_order = inherited if inherited else []

    # ---- identity / navigation ---------------------------------------------------------
    def uuid_str(self):
        return str(self.uuid)

    def copy(self):
        """Fresh node (new uuid, no children) over copies of this board and tile list."""
        return State(self.board, self.tiles, parent=self.parent)

    def child_with_biggest_score(self):
        return max(self.children, key=sort_key)

    def centralize_node_with_children(self):
        """Along the path of expanded nodes, move the expanded child of every node to the
        middle slot of its sibling list (engine/state.py:116-136; tree drawing only)."""
        node = self
        while node.children:
            expanded = [i for i, child in enumerate(node.children) if child.children]
            if not expanded:
                return
            src = expanded[-1]
            mid = (len(node.children) - 1) // 2
            follow = node.children[src]
            node.children[mid], node.children[src] = node.children[src], node.children[mid]
            node = follow

    # ---- exports ------------------------------------------------------------------------
    def to_json(self):
        return OrderedDict([
            ('tiles', self.tiles),
            ('board', self.board.tolist()),
            ('tile_placed', self.tile_placed),
            ('score', self.score if self.score else 0),
            ('name', self.uuid_str()),
            ('children', []),
        ])

    def render_to_json(self):
        """Nested {'name', 'board', 'tiles', 'score', 'tile_placed', 'children': [...]}."""
        out = self.to_json()
        out['children'] = [child.render_to_json() for child in self.children]
        return out

    def render_children(self, only_ids=False):
        """({label: {child label: {...}}}, {uuid: node}) with uuid or text labels."""
        index = {}

        def label(node):
            return node.uuid_str() if only_ids else str(node)

        def walk(node):
            index[node.uuid_str()] = node
            return OrderedDict((label(child), walk(child)) for child in node.children)

        return {label(self): walk(self)}, index

    def __repr__(self, short=False):
        if short:
            return str(self.tiles)
        shown = [list(t) for t in self.tiles]
        if len(shown) > 6:
            shown = '%s(...)' % (shown[:6],)
        tail = self.board if len(self.tiles) <= 4 else ''
        return '({}) Remaining tiles: {}, Tile placed: {}. Tiles: {}. Sim. depth:({}) {}'.format(
            self.uuid, len(self.tiles) / ORIENTATIONS, self.tile_placed, shown, self.score, tail)

    __str__ = __repr__


def render_to_dict(node, tree=None, all_nodes=None, only_ids=False):
    """Function form kept for callers of the reference's module-level helper."""
    nested, index = node.render_children(only_ids=only_ids)
    if all_nodes is not None:
        all_nodes.update(index)
        index = all_nodes
    return next(iter(nested.values())), index


def render_to_json(node, tree=None, i=0, all_nodes=None, only_ids=False):
    return node.render_to_json()
