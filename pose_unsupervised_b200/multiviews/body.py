"""Skeleton trees for the pictorial model, behind the name of lib/multiviews/body.py:11-57.

``HumanBody()`` is the reference's 16-joint MPII-ordered tree (root 6).  The
reference has no 17-joint tree although BASELINE.json's configs use 17 joints;
``HumanBody.h36m17()`` supplies the anatomical tree over the H36M joint order of
lib/dataset/multiview_h36m_compatible.py:26-44.  ``tree_arrays()`` is what the RPSM
kernel consumes.
"""
import numpy as np

_MPII16 = (['rank', 'rkne', 'rhip', 'lhip', 'lkne', 'lank', 'root', 'thorax', 'upper neck',
            'head top', 'rwri', 'relb', 'rsho', 'lsho', 'lelb', 'lwri'],
           [[], [0], [1], [4], [5], [], [2, 3, 7], [8, 12, 13], [9], [], [], [10], [11], [14], [15], []],
           6)
_H36M17 = (['root', 'rhip', 'rkne', 'rank', 'lhip', 'lkne', 'lank', 'belly', 'neck', 'nose', 'head',
            'lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri'],
           [[1, 4, 7], [2], [3], [], [5], [6], [], [8], [9, 11, 14], [10], [], [12], [13], [], [15],
            [16], []],
           0)


class HumanBody(object):

    def __init__(self, joint_names=None, children=None, root_idx=None):
        if joint_names is None:
            joint_names, children, root_idx = _MPII16
        self.root_idx = root_idx
        self.skeleton = [{'idx': i, 'name': n, 'children': list(ch)}
                         for i, (n, ch) in enumerate(zip(joint_names, children))]
        self.skeleton_sorted_by_level = self.sort_skeleton_by_level(self.skeleton)

    @classmethod
    def h36m17(cls):
        return cls(*_H36M17)

    def sort_skeleton_by_level(self, skeleton):
        """Breadth-first levels from the root, deepest joints first (body.py:39-57)."""
        level = np.zeros(len(skeleton))
        pending = [skeleton[self.root_idx]]
        while pending:
            node = pending.pop(0)
            for ch in node['children']:
                skeleton[ch]['parent'] = node['idx']
                level[ch] = level[node['idx']] + 1
                pending.append(skeleton[ch])
        by_depth = np.argsort(level, kind='stable')[::-1]
        for i in by_depth:
            skeleton[i]['level'] = level[i]
        return [skeleton[i] for i in by_depth]

    def edges(self):
        """(parent, child) in the order pictorial.py iterates them (skeleton, then children)."""
        return [(n['idx'], c) for n in self.skeleton for c in n['children']]

    def tree_arrays(self):
        """(edges [E,2] int32, order [J] int32 children-before-parents, root_idx)."""
        edges = np.array(self.edges(), dtype=np.int32).reshape(-1, 2)
        order = np.array([n['idx'] for n in self.skeleton_sorted_by_level], dtype=np.int32)
        if len(edges) != len(self.skeleton) - 1:
            raise ValueError('skeleton is not a tree')
        return edges, order, int(self.root_idx)
