"""Pass-through head of the supervised ``*-orig`` configs (reference ``src/heads/NoOpHead.py``): picks
``(ground_truth, network_output, delta_gt, delta_hat)`` out of the batch dict for an external ``torch.nn`` loss.
Same kwargs (TARGET_GEN, LEARNING_KEYS), same return values.

``predict_homography`` for '4_points' runs the 4-point DLT on K1 (``four_point_to_homography``); for 'all_points'
(dense perspective field, Zeng-orig) it is the reference's CPU post-processing -- ``cv2.findHomography(RANSAC, 10)``
over all P*P correspondences per sample (reference :64-109) -- kept on OpenCV so that evaluation reports the same
corner errors; it is an eval-only path and not part of the training step.
"""
import numpy as np
import torch
import torch.nn as nn

from ..data.utils import four_point_to_homography


class Model(nn.Module):

    def __init__(self, backbone, **kwargs):
        super().__init__()
        self.target_gen = kwargs['TARGET_GEN']          # '4_points' | 'all_points'
        self.learning_keys = kwargs['LEARNING_KEYS']    # ground_truth, network_output, delta_gt, delta_hat
        assert self.target_gen in ('4_points', 'all_points'), 'I didnt understand that!'

    def forward(self, data):
        ret = [data[key] for key in self.learning_keys[:-1]]
        last = data[self.learning_keys[-1]]
        if self.target_gen == 'all_points':
            # the field's values at the four patch corners, clockwise from the top left (reference :33-49)
            h, w = last.shape[-2:]
            ys = torch.tensor([0, 0, h - 1, h - 1], device=last.device)
            xs = torch.tensor([0, w - 1, w - 1, 0], device=last.device)
            last = last[:, :2, ys, xs].transpose(1, 2)
        ret.append(last)
        return ret

    def predict_homography(self, data):
        if self.target_gen == '4_points':
            assert 'corners' in data, 'How to handle it?'
            delta_hat = data[self.learning_keys[3]]
            return delta_hat, four_point_to_homography(corners=data['corners'], deltas=delta_hat, crop=False)
        return self._postprocess(data[self.learning_keys[1]])

    @staticmethod
    def _postprocess(perspective_field):
        """[B,2,h,w] field -> (delta [B,4,2], H [B,3,3]) as numpy, by RANSAC over every pixel's correspondence"""
        import cv2
        if torch.is_tensor(perspective_field):
            perspective_field = perspective_field.detach().cpu().numpy()
        b, _, h, w = perspective_field.shape
        ys, xs = np.mgrid[0:h, 0:w]
        coords = np.stack((xs.reshape(-1), ys.reshape(-1)), axis=-1)                       # [h*w, 2] (x, y)
        moved = coords[None] + perspective_field.reshape(b, 2, -1).transpose(0, 2, 1)     # [b, h*w, 2]
        four = [[0, 0], [w, 0], [w, h], [0, h]]
        hs, deltas = [], []
        for i in range(b):
            hom = cv2.findHomography(np.float32(coords), np.float32(moved[i]), cv2.RANSAC, 10)[0]
            if hom is None:     # no consensus (an untrained network's field): the reference would raise here
                hom = cv2.findHomography(np.float32(coords), np.float32(moved[i]), 0)[0]
            if hom is None:
                hom = np.eye(3)
            deltas.append(cv2.perspectiveTransform(np.asarray([four], dtype=np.float32), hom).squeeze() - four)
            hs.append(hom)
        return np.array(deltas), np.array(hs)
