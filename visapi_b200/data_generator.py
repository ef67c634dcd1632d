"""Instance generator and reader (reference: engine/data_generator.py:24-135).

Host-side Python: a guillotine splitter that cuts the w x h bin n-1 times, and the
CSV reader of the reference's puzzle file.  The random draws are made in the same
order and from the same generators as the reference (``random.randint`` for the bin
to split, ``numpy.random.randint`` for axis and cut), so a given (``random.seed``,
numpy seed) pair yields the same instance on both sides.
"""
import csv
import json
import random
from collections import OrderedDict, defaultdict, namedtuple

import numpy as np

InstanceFromFile = namedtuple('InstanceFromFile', ['bins', 'id', 'n_tiles_placed'])

ORIENTATIONS = 2


class DataGenerator(object):

    def __init__(self, w=None, h=None):
        self.w = w
        self.h = h
        self.instances_from_file = defaultdict(lambda: OrderedDict())

    # ---- generation -----------------------------------------------------------------
    def _split_bin(self, _bin, axis, value):
        """Cut [a, b, (x, y)] at ``value`` along ``axis`` into two bins (:316-326)."""
        a, b, (x, y) = _bin
        if axis == 0:
            return [[value, b, (x, y)], [a - value, b, (x + value, y)]]
        return [[a, value, (x, y)], [a, b - value, (x, y + value)]]

    def gen_instance_visual(self, n, w, h, dimensions=2, seed=None):
        """n bins [a, b, (x, y)] tiling the w x h bin exactly (:24-64)."""
        self.w, self.h = w, h
        if seed is not None:
            np.random.seed(seed)
        bins = [[w, h, (0, 0)]]
        while len(bins) < n:
            k = random.randint(0, len(bins) - 1)
            target = bins[k]
            axis = np.random.randint(0, 2, size=1)[0]
            if target[axis] <= 1:
                continue                      # cannot be cut along this axis
            cut = int(np.random.randint(1, target[axis], size=1)[0])
            first, second = self._split_bin(target, axis, cut)
            bins[k:k + 1] = [second, first]
        return bins

    def _transform_instance_visual_to_np_array(self, bins, dimensions=2):
        tiles = np.array([b[:dimensions] for b in bins])
        solution = sorted(bins, key=lambda b: (b[2][0], b[2][1]))
        return tiles, solution

    def gen_instance(self, n, w, h, dimensions=2, seed=0):
        tiles, solution = self._transform_instance_visual_to_np_array(
            self.gen_instance_visual(n, w, h, seed=seed), dimensions=dimensions)
        return np.array(tiles), solution

    def gen_tiles_and_board(self, n, w, h, dimensions=2, seed=0, order_tiles=False, from_file=False):
        """(tile list with both orientations sorted by (x[1], x[0]), zero board w x h) (:75-98)."""
        if from_file:
            visual = self.gen_instance_from_file(n, w, h)
        else:
            visual = self.gen_instance_visual(n, w, h, seed=seed)
        return self.transform_instance_visual_to_tiles_and_board(
            w, h, visual, dimensions=dimensions, order_tiles=order_tiles)

    def transform_instance_visual_to_tiles_and_board(self, w, h, instance_visual, dimensions=2,
                                                     order_tiles=False):
        tiles, _ = self._transform_instance_visual_to_np_array(instance_visual, dimensions=dimensions)
        both = []
        for t in tiles:
            both.append(tuple(t))
            both.append((t[1], t[0]))
        both = sorted(both, key=lambda x: (x[1], x[0]))     # the reference always sorts (:93)
        return both, np.zeros((w, h))

    # ---- reading --------------------------------------------------------------------
    def gen_instance_from_file(self, n, w, h):
        return self.read_instances()[n][w, h][0].bins

    def read_instances(self, path='puzzles_n20_with_solutions_2.csv'):
        """{num_tiles: {(board_width, board_height): [InstanceFromFile, ...]}} (:117-125)."""
        with open(path, newline='') as f:
            for row in csv.DictReader(f, delimiter=',', quotechar='"'):
                bins = DataGenerator.x_y_str_to_bins_format(row['tiles'])
                key = (int(row['board_width']), int(row['board_height']))
                self.instances_from_file[int(row['num_tiles'])].setdefault(key, []).append(
                    InstanceFromFile(bins=bins, id=row['id'], n_tiles_placed=row['tiles_placed']))
        return self.instances_from_file

    @staticmethod
    def x_y_str_to_bins_format(x_y_string):
        """'[{"X": a, "Y": b}, ...]' -> [[a, b, (0, 0)], ...]; parsed as JSON, not eval (:127-135)."""
        return [[el['X'], el['Y'], (0, 0)] for el in json.loads(x_y_string)]
