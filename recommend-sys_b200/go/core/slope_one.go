package core

// slope_one.go — replaces core/slope_one.go of the reference (SURVEY.md §8 f-2): the item x item
// deviation matrix is computed on the device through the same C ABI as the KNN similarities
// (rs_knn_params.sim = RS_SIM_SLOPE_ONE; the co-rated sums count, sum r_i, sum r_j are integer
// contractions on the tensor cores), Predict is a batch gather.  Same type and constructor names
// as the reference; `dev` is materialised on demand.

/*
#include "rs_knn.h"
*/
import "C"

import "runtime"

type SlopeOne struct {
	Base
	globalMean float64
	dev        *deviceKNN // unexported: the gob round-trip of Copy never sees the handle
	nItems     int
}

func NewSlopeOne(params Parameters) *SlopeOne { // core/slope_one.go:16-20
	return &SlopeOne{Base: Base{Params: params}}
}

func (s *SlopeOne) Close() {
	if s.dev != nil {
		s.dev.close()
		s.dev = nil
	}
}

func (s *SlopeOne) Fit(trainSet TrainSet) { // core/slope_one.go:47-93
	s.Data = trainSet
	s.globalMean = trainSet.GlobalMean
	n := trainSet.Length()
	items, users := make([]int32, n), make([]int32, n)
	for i := 0; i < n; i++ {
		items[i] = int32(trainSet.ConvertItemID(trainSet.Items[i]))
		users[i] = int32(trainSet.ConvertUserID(trainSet.Users[i]))
	}
	s.Close()
	s.dev = newDeviceSlopeOne(s.Params)
	s.dev.fit(items, users, trainSet.Ratings, trainSet.ItemCount, trainSet.UserCount, trainSet.GlobalMean, nil, nil, 0)
	s.nItems = trainSet.ItemCount
}

// PredictBatch implements BatchPredictor (data_batch.go): the whole test set in one cgo call.
func (s *SlopeOne) PredictBatch(userIDs, itemIDs []int) []float64 {
	left, right := make([]int32, len(userIDs)), make([]int32, len(userIDs))
	for i := range userIDs {
		left[i] = int32(s.Data.ConvertItemID(itemIDs[i]))   // -1 = newID (core/data.go:129)
		right[i] = int32(s.Data.ConvertUserID(userIDs[i]))
	}
	return s.dev.predictBatch(left, right)
}

func (s *SlopeOne) Predict(userId int, itemId int) float64 { // core/slope_one.go:22-45
	return s.PredictBatch([]int{userId}, []int{itemId})[0]
}

// Dev materialises rows [row0, row0+nrows) of the deviation matrix (core/slope_one.go:13).
func (s *SlopeOne) Dev(row0, nrows int) []float64 { return s.dev.simsRows(row0, nrows, s.nItems) }

func newDeviceSlopeOne(p Parameters) *deviceKNN {
	var cp C.rs_knn_params
	check(C.rs_knn_params_default(&cp))
	cp.sim = C.RS_SIM_SLOPE_ONE
	cp.device = C.int32_t(p.GetInt("device", -1))
	d := &deviceKNN{}
	check(C.rs_knn_create(&cp, &d.h))
	// CrossValidate's copies (core/eval.go:29-35) are never Closed: release the device memory with the object
	runtime.SetFinalizer(d, func(d *deviceKNN) { d.close() })
	return d
}
