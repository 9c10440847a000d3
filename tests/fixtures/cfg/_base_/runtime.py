optimizer = dict(type='SGD', lr=0.02, momentum=0.9, weight_decay=0.0001)
dist_params = dict(backend='nccl')
log_level = 'INFO'
