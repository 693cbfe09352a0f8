"""Literal Python model of the reference's incremental tree clustering.  TEST INFRASTRUCTURE ONLY.

Restates the live code of src/tree.rs (dead in the reference: `// mod tree;` at src/main.rs:15,
and it calls Protein accessors that do not exist).  The two missing accessors are defined as in
SURVEY.md §8c: get_five_hash_map() = presence bit-array over the repeated-k-mer ids,
get_five_hash() = the id list.  Sets stand in for the (bit-array, index-list) pairs; only set
contents and the loop order of `balance` influence the result.

    Node::new_leaf ............ src/tree.rs:64-106
    Node::clone_and_clean ..... src/tree.rs:151-177
    Node::balance ............. src/tree.rs:179-265
    Node::add_child ........... src/tree.rs:267-385
    Tree::new / add_protein ... src/tree.rs:519-536

PARITY: unpinned by the reference (the module never compiled); this model is the specification
the C++ host implementation (csrc/tree.cpp) is tested against.
"""
from __future__ import annotations


class Node:
    __slots__ = ("children", "u", "c", "protein")

    def __init__(self, children, u, c, protein):
        self.children, self.u, self.c, self.protein = children, u, c, protein


def new_leaf(protein: int, ids) -> Node:
    s = frozenset(int(x) for x in ids)
    return Node([], s, s, protein)


def balance(curr: Node, log=None):
    best = (0, 0, 0)
    mn = None
    ch = curr.children
    for i in range(1, len(ch)):
        ci = ch[i].c
        for j in range(i):
            s = len(ci & ch[j].c)
            if s > best[0]:
                best = (s, i, j)
            if mn is None or mn > s:
                mn = s
    if best[0] > mn:
        if log is not None:
            log.append("Merging")
        one, two = ch[best[1]], ch[best[2]]
        if len(one.children) < len(two.children):
            ch.pop(best[2])
            add_child(one, two, log)
        else:
            ch.pop(best[1])
            add_child(two, one, log)


def add_child(curr: Node, child: Node, log=None):
    if not curr.children:
        cloned = Node([], curr.u, curr.c, curr.protein)      # clone_and_clean
        curr.protein = None
        curr.u = cloned.u | child.u
        curr.c = cloned.c & child.c
        curr.children = [cloned]
        if not child.children:
            curr.children.append(child)
        else:
            curr.children.extend(child.children)
    else:
        common = bool(curr.u & child.u)
        curr.u = curr.u | child.u
        curr.c = curr.c & child.c
        curr.children.append(child)
        if common:
            balance(curr, log)
        elif log is not None:
            log.append("No kmers in common")


class Tree:
    def __init__(self, protein: int, ids):
        self.root = new_leaf(protein, ids)
        self.log = []

    def add_protein(self, protein: int, ids):
        add_child(self.root, new_leaf(protein, ids), self.log)


def build_tree(id_rows) -> Tree:
    """id_rows[p] = repeated-k-mer ids of protein p; proteins are added in index order."""
    t = Tree(0, id_rows[0])
    for p in range(1, len(id_rows)):
        t.add_protein(p, id_rows[p])
    return t


def leaves(node: Node):
    if node.protein is not None and not node.children:
        return [node.protein]
    out = []
    for ch in node.children:
        out.extend(leaves(ch))
    return out


def nested(node: Node):
    """canonical nested-list form of the tree (children in order)"""
    if not node.children:
        return node.protein
    return [nested(ch) for ch in node.children]


def clusters(tree: Tree):
    """the top-level clusters = the root's children, as lists of proteins"""
    if not tree.root.children:
        return [[tree.root.protein]]
    return [leaves(ch) for ch in tree.root.children]
