"""Host-side mirror of the slice of `Poll` / `PollProvider` that drives the hot
path (pallet/src/poll/provider.rs, pallet/src/poll/state.rs:14-67):

    PollState::new(registration_depth, interaction_depth)      state.rs:42-67
    register_participant(public_key, timestamp)                provider.rs:218-241
    consume_interaction(public_key, data)                      provider.rs:243-287
    merge_registrations() / merge_interactions()               provider.rs:289-327
    prepare_public_inputs(coordinator, new_commitment)         provider.rs:141-215
    get_voting_period_end()                                    provider.rs:355-358
    registration_limit_reached / interaction_limit_reached     provider.rs:329-341

Leaves are hashed in bulk on the GPU when they are needed (at merge time, or
when `leaves()` is asked for), not one extrinsic at a time.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .context import Context, get_context
from .errors import MerkleTreeError
from .leaves import interaction_leaves, registration_leaves, replay_interactions, replay_registrations
from .tree import (PollStateTree, merge_interactions as _merge_interactions,
                   merge_registrations as _merge_registrations, new_interaction_tree, new_registration_tree)


@dataclass
class Commitment:                       # coordinator.rs, the fields this path writes
    process: Tuple[int, bytes] = (0, bytes(32))
    tally: Tuple[int, bytes] = (0, bytes(32))
    expected_process: int = 0
    expected_tally: int = 0


@dataclass
class PollConfig:                       # config.rs, the fields this path reads
    registration_depth: int
    interaction_depth: int
    process_subtree_depth: int
    tally_subtree_depth: int
    signup_period: int = 0
    voting_period: int = 0


class Poll:
    def __init__(self, config: PollConfig, ctx: Optional[Context] = None, created_at: int = 0):
        self.ctx = ctx or get_context()
        self.config = config
        self.created_at = created_at
        self.registrations: PollStateTree = new_registration_tree(config.registration_depth, self.ctx)
        self.interactions: PollStateTree = new_interaction_tree(config.interaction_depth, self.ctx)
        self.commitment = Commitment()
        self._reg_pk: List[bytes] = []
        self._reg_ts: List[int] = []
        self._msg_pk: List[bytes] = []
        self._msg_data: List[bytes] = []

    # provider.rs:329-341
    def registration_limit_reached(self) -> bool:
        return self.registrations.count + len(self._reg_pk) >= 2 ** self.config.registration_depth - 1

    def interaction_limit_reached(self) -> bool:
        return self.interactions.count + len(self._msg_pk) >= 5 ** self.config.interaction_depth

    def register_participant(self, public_key: Tuple[bytes, bytes], timestamp: int) -> int:
        """Returns the registration count, as the reference does."""
        if self.registrations.root is not None:
            raise MerkleTreeError("TreeAlreadyFull")
        self._reg_pk.append(bytes(public_key[0]) + bytes(public_key[1]))
        self._reg_ts.append(int(timestamp))
        return self.registrations.count + len(self._reg_pk)

    def consume_interaction(self, public_key: Tuple[bytes, bytes], data) -> int:
        if self.interactions.root is not None:
            raise MerkleTreeError("TreeAlreadyFull")
        d = b"".join(bytes(x) for x in data)
        if len(d) != 320:
            raise ValueError("PollInteractionData is 10 x 32 bytes")
        self._msg_pk.append(bytes(public_key[0]) + bytes(public_key[1]))
        self._msg_data.append(d)
        return self.interactions.count + len(self._msg_pk)

    def _flush(self):
        if self._reg_pk:
            lv = registration_leaves(np.frombuffer(b"".join(self._reg_pk), dtype=np.uint8),
                                     np.array(self._reg_ts, dtype=np.uint64), self.ctx)
            self._reg_pk, self._reg_ts = [], []
            self.registrations.extend(lv)
        if self._msg_pk:
            lv = interaction_leaves(np.frombuffer(b"".join(self._msg_pk), dtype=np.uint8),
                                    np.frombuffer(b"".join(self._msg_data), dtype=np.uint8), self.ctx)
            self._msg_pk, self._msg_data = [], []
            self.interactions.extend(lv)

    def _untouched(self, tree: PollStateTree) -> bool:
        """Nothing hashed in yet: the whole poll can be replayed raw rows -> root on the device."""
        return tree.root is None and tree._fresh and tree._pending == 0

    def merge_registrations(self) -> "Poll":
        if self._reg_pk and self._untouched(self.registrations):
            pk, ts = np.frombuffer(b"".join(self._reg_pk), dtype=np.uint8), np.array(self._reg_ts, dtype=np.uint64)
            self.registrations, c, _, _ = replay_registrations(self.config.registration_depth, pk, ts, self.ctx)
            self._reg_pk, self._reg_ts = [], []
            self.commitment.process = (0, c)
            return self
        self._flush()
        self.registrations, c = _merge_registrations(self.registrations)
        self.commitment.process = (0, c)
        return self

    def merge_interactions(self) -> "Poll":
        if self._msg_pk and self._untouched(self.interactions):
            pk = np.frombuffer(b"".join(self._msg_pk), dtype=np.uint8)
            data = np.frombuffer(b"".join(self._msg_data), dtype=np.uint8)
            self._msg_pk, self._msg_data = [], []
            self._flush()                                       # registrations still pending are hashed the usual way
            self.interactions, ep, et, _, _ = replay_interactions(
                self.config.interaction_depth, pk, data, self.registrations.count, self.config.process_subtree_depth,
                self.config.tally_subtree_depth, self.ctx)
            self.commitment.expected_process, self.commitment.expected_tally = ep, et
            return self
        self._flush()
        self.interactions, ep, et = _merge_interactions(self.interactions, self.registrations.count,
                                                        self.config.process_subtree_depth,
                                                        self.config.tally_subtree_depth)
        self.commitment.expected_process, self.commitment.expected_tally = ep, et
        return self

    def get_voting_period_end(self) -> int:                       # provider.rs:355-358
        return self.created_at + self.config.signup_period + self.config.voting_period

    def prepare_public_inputs(self, coordinator_public_key: Tuple[bytes, bytes], new_commitment: bytes):
        """provider.rs:141-215 without the verify key: ("process" | "tally", the public
        inputs as integers, the commitment the proof would install) or None."""
        from .hasher import MODULUS, Poseidon
        fr = lambda b: int.from_bytes(bytes(b), "big") % MODULUS      # Fr::from_be_bytes_mod_order
        reg, it, c = self.registrations, self.interactions, self.commitment
        message_batch_size = it.arity ** self.config.process_subtree_depth
        current_batch_index = it.count
        if current_batch_index > 0:
            r = it.count % message_batch_size
            current_batch_index -= message_batch_size if r == 0 else r
        proof_index = c.process[0]
        index_offset = proof_index * message_batch_size
        if index_offset <= current_batch_index:
            coord_hash = Poseidon.new_circom(2, self.ctx).hash([fr(coordinator_public_key[0]),
                                                                fr(coordinator_public_key[1])])
            if it.root is None:
                return None
            current_batch_index -= index_offset
            end_batch_index = min(current_batch_index + message_batch_size, it.count)
            inputs = [reg.count + 1, self.get_voting_period_end(), fr(it.root), reg.depth, end_batch_index,
                      current_batch_index, coord_hash, fr(c.process[1]), fr(new_commitment)]
            nxt = Commitment(process=(proof_index + 1, bytes(new_commitment)), tally=c.tally,
                             expected_process=c.expected_process, expected_tally=c.expected_tally)
            return "process", inputs, nxt
        proof_index = c.tally[0]
        batch_size = reg.arity ** self.config.tally_subtree_depth
        current_batch_index = proof_index * batch_size
        if current_batch_index >= reg.count + 1:
            return None
        inputs = [fr(c.process[1]), fr(c.tally[1]), fr(new_commitment), current_batch_index, reg.count + 1]
        nxt = Commitment(process=c.process, tally=(proof_index + 1, bytes(new_commitment)),
                         expected_process=c.expected_process, expected_tally=c.expected_tally)
        return "tally", inputs, nxt
