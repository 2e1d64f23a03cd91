"""Oracle: skeleton trees for the pictorial model (TEST INFRASTRUCTURE).

``HumanBody()`` restates lib/multiviews/body.py:11-57 (16 MPII-ordered joints,
root 6).  The reference has no 17-joint tree; ``h36m17()`` is the anatomical
tree over the H36M joint order of
lib/dataset/multiview_h36m_compatible.py:26-44 that BASELINE.json's 17-joint
configs need.
"""
import numpy as np

MPII16_NAMES = ['rank', 'rkne', 'rhip', 'lhip', 'lkne', 'lank', 'root', 'thorax',
                'upper neck', 'head top', 'rwri', 'relb', 'rsho', 'lsho', 'lelb', 'lwri']
MPII16_CHILDREN = [[], [0], [1], [4], [5], [], [2, 3, 7], [8, 12, 13], [9], [],
                   [], [10], [11], [14], [15], []]
MPII16_ROOT = 6

H36M17_NAMES = ['root', 'rhip', 'rkne', 'rank', 'lhip', 'lkne', 'lank', 'belly', 'neck',
                'nose', 'head', 'lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri']
H36M17_CHILDREN = [[1, 4, 7], [2], [3], [], [5], [6], [], [8], [9, 11, 14], [10], [],
                   [12], [13], [], [15], [16], []]
H36M17_ROOT = 0


class HumanBody(object):
    """Tree with the attributes lib/multiviews/pictorial.py reads off the body."""

    def __init__(self, names=None, children=None, root_idx=None):
        names = MPII16_NAMES if names is None else names
        children = MPII16_CHILDREN if children is None else children
        self.root_idx = MPII16_ROOT if root_idx is None else root_idx
        self.skeleton = [{'idx': i, 'name': names[i], 'children': list(children[i])}
                         for i in range(len(names))]
        self.skeleton_sorted_by_level = self._by_level_desc()

    def _by_level_desc(self):
        # lib/multiviews/body.py:39-57 -- BFS levels, deepest first
        sk = self.skeleton
        level = np.zeros(len(sk))
        fifo = [sk[self.root_idx]]
        while fifo:
            cur = fifo.pop(0)
            for ch in cur['children']:
                sk[ch]['parent'] = cur['idx']
                level[ch] = level[cur['idx']] + 1
                fifo.append(sk[ch])
        order = np.argsort(level)[::-1]
        for i in order:
            sk[i]['level'] = level[i]
        return [sk[i] for i in order]

    def edges(self):
        """(parent, child) pairs in the reference's iteration order (skeleton, then children)."""
        return [(n['idx'], c) for n in self.skeleton for c in n['children']]


def h36m17():
    return HumanBody(H36M17_NAMES, H36M17_CHILDREN, H36M17_ROOT)
