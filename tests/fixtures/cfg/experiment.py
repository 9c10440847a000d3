_base_ = ['./_base_/model.py', './_base_/runtime.py']
model = dict(roi_head=dict(bbox_head=dict(num_classes=3)))
optimizer = dict(type='SGD', lr=0.001, momentum=0.9, weight_decay=0.0005)
