"""Search-tree node (reference: engine/state.py:63-153) and its JSON / dict renderers."""
from collections import OrderedDict

import numpy as np

ORIENTATIONS = 2


class _Counter(object):
    def __init__(self):
        self.i = 0

    def uuid(self):
        self.i += 1
        return self.i


_uuid = _Counter()


def sort_key(x):
    return x.score


class State(object):
    """board is copied (np.copy) and tiles is sliced on construction, as the reference
    does (state.py:66-67), so a State never aliases its parent's board."""

    def __init__(self, board, tiles, parent=None, solution_tiles_order=None):
        self.board = np.copy(board)
        self.tiles = tiles[:]
        self.parent = parent
        self.uuid = _uuid.uuid()
        self.children = []
        self.score = None
        self.tile_placed = None
        if parent and parent.solution_tiles_order:
            self.solution_tiles_order = parent.solution_tiles_order
        else:
            self.solution_tiles_order = []

    def uuid_str(self):
        return '%s' % self.uuid

    def copy(self):
        return State(self.board, self.tiles, parent=self.parent)

    def child_with_biggest_score(self):
        return sorted(self.children, key=sort_key, reverse=True)[0]

    def to_json(self):
        return {
            'tiles': self.tiles,
            'board': self.board.tolist(),
            'tile_placed': self.tile_placed,
            'score': self.score or 0,
            'name': self.uuid_str(),
            'children': [],
        }

    def render_to_json(self):
        return render_to_json(self)

    def render_children(self, only_ids=False):
        ret, all_nodes = render_to_dict(self, all_nodes={}, only_ids=only_ids)
        self_str = self.uuid_str() if only_ids else str(self)
        all_nodes[self.uuid_str()] = self
        return {self_str: ret}, all_nodes

    def centralize_node_with_children(self):
        """Moves the child that has children to the middle slot, level by level
        (state.py:116-136; used for the tree visualisation)."""
        _state = self
        while _state.children:
            index_with_children = None
            for i, child in enumerate(_state.children):
                if child.children:
                    index_with_children = i
            if index_with_children is None:
                break
            n = len(_state.children)
            middle = n // 2 - 1 if n % 2 == 0 else n // 2
            nxt = _state.children[index_with_children]
            kids = _state.children
            kids[middle], kids[index_with_children] = kids[index_with_children], kids[middle]
            _state = nxt

    def __str__(self):
        return self.__repr__()

    def __repr__(self, short=False):
        if short:
            return '%s' % (self.tiles,)
        output_list = [list(x) for x in self.tiles]
        if len(output_list) > 6:
            output_list = str(output_list[:6]) + '(...)'
        output_board = ''
        if len(self.tiles) <= 4:
            output_board = self.board
        return ('(%s) Remaining tiles: %s, Tile placed: %s. Tiles: %s. Sim. depth:(%s) %s' % (
            self.uuid, len(self.tiles) / ORIENTATIONS, self.tile_placed, output_list, self.score,
            output_board))


def render_to_dict(node, tree=None, all_nodes=None, only_ids=False):
    """Nested OrderedDict of the subtree plus {uuid: node} (state.py:16-43)."""
    if all_nodes is None:
        all_nodes = {}
    all_nodes[node.uuid_str()] = node
    if not node.children:
        return {}, all_nodes
    sub = OrderedDict()
    for child in node.children:
        key = child.uuid_str() if only_ids else str(child)
        sub[key], all_nodes = render_to_dict(child, None, all_nodes, only_ids=only_ids)
    return sub, all_nodes


def render_to_json(node):
    """{..., 'children': [...]} recursively (state.py:45-61)."""
    out = node.to_json()
    for child in node.children:
        out['children'].append(render_to_json(child))
    return out
